"""runtime.TowerPipeline / DevicePrefetcher(lookahead=1): running the frozen towers of batch i + 1 under the tail of batch i must
not change a single loss or parameter — same kernels, same inputs, different streams and buffer slots."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(seed=0):
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import parallel_config
    cfg = parallel_config("tiny")
    cfg["model_settings"]["parallel_branch"]["transformer_args"]["dropout"] = 0.0  # dropout state advances per step either way; keep the comparison exact
    torch.manual_seed(seed)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg)).to(DEV).train()
    opts, _ = model.configure_optimizers()
    return model, opts[0]


def _host_batches(n, B=6, seed=3):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        lens = torch.randint(4000, 6001, (B,), generator=g)
        lens[0] = 6000
        out.append({"wav": (0.1 * torch.randn(B, 6000, generator=g)).pin_memory(), "wav_len": lens.pin_memory(),
                    "image": torch.randn(B, 3, 32, 32, generator=g).pin_memory(), "id": (torch.arange(B) + i * B).pin_memory()})
    return out


def _train(model, opt, batches):
    losses = []
    for b in batches:
        loss = model.training_step_end(model.training_step(b))["loss"]
        opt.zero_grad()
        loss.backward()
        model.on_after_backward()
        opt.step()
        losses.append(loss.detach().clone())
    torch.cuda.synchronize()
    return [float(x) for x in losses]


@pytest.mark.parametrize("graphs", [False, True])
def test_pipelined_steps_equal_the_plain_loop(graphs):
    from speechclip_b200 import engine
    from speechclip_b200.runtime import DevicePrefetcher, TowerPipeline
    old = engine.GRAPHS
    engine.GRAPHS = graphs
    try:
        host = _host_batches(7)
        plain, opt_p = _model()
        ref = _train(plain, opt_p, DevicePrefetcher(host, torch.device(DEV)))
        piped, opt_q = _model()
        got = _train(piped, opt_q, TowerPipeline(piped).iterate(DevicePrefetcher(host, torch.device(DEV), lookahead=1)))
        assert got == pytest.approx(ref, rel=1e-6, abs=1e-7), (got, ref)   # the same kernels on the same inputs ...
        for (n, a), (_, b) in zip(plain.named_parameters(), piped.named_parameters()):
            if a.requires_grad:  # ... up to the summation order of the atomics in the bias / column-sum reductions
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), n
    finally:
        engine.GRAPHS = old


def test_pipeline_needs_a_lookahead_slot():
    from speechclip_b200.runtime import DevicePrefetcher, TowerPipeline
    model, _ = _model()
    with pytest.raises(ValueError, match="lookahead=1"):
        next(iter(TowerPipeline(model).iterate(DevicePrefetcher(_host_batches(2), torch.device(DEV)))))


def test_prefetcher_slots_are_not_overwritten_while_held():
    """With lookahead=1 the consumer holds two batches; both must still carry their own data after the third was staged."""
    from speechclip_b200.runtime import DevicePrefetcher
    host = [{"x": torch.full((1 << 20,), float(i)).pin_memory()} for i in range(6)]
    it = iter(DevicePrefetcher(host, torch.device(DEV), lookahead=1))
    held = [next(it), next(it)]
    for i in range(2, 6):
        sums = [float(h["x"].mean()) for h in held]
        assert sums == [float(i - 2), float(i - 1)], sums
        held = [held[1], next(it)]
