"""Regression tests for defects found in review: each reproduces the failing sequence on the GPU and checks
the result against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-8)).item()


def _branch_pair(d=64, heads=4, ffn=128, out=32, seed=0):
    from avssl.base import OrderedNamespace
    from avssl.model.kwClip import KW_ParallelBranch
    from oracle import speechclip as osc
    from speechclip_b200.configs import parallel_config
    cfg = parallel_config("tiny")
    cfg["model_settings"]["parallel_branch"]["transformer_args"].update(d_model=d, nhead=heads, dim_feedforward=ffn)
    torch.manual_seed(seed)
    mine = KW_ParallelBranch(OrderedNamespace(cfg), d, out)
    ref = osc.ParallelBranch(d, out, n_layers=1, nhead=heads, dim_feedforward=ffn).eval()
    ref.load_state_dict(mine.state_dict())
    return mine.to(DEV).eval(), ref


def _branch_grads(mine, ref, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, T, mine.audio_dim, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    w = torch.randn(B, mine.out_dim, generator=g)
    for m in (mine, ref):
        m.zero_grad()
    (mine(feat.to(DEV), lens.to(DEV)) * w.to(DEV)).sum().backward()
    (ref(feat, lens) * w).sum().backward()
    name = "self_att.model.layers.0.self_attn.in_proj_weight"
    return dict(mine.named_parameters())[name].grad.cpu(), dict(ref.named_parameters())[name].grad


def test_split_k_wgrad_after_a_longer_batch_in_the_same_bucket():
    """the split-K K/V wgrad contracts over ldt = ceil(M / 256) * 256 columns of two workspace buffers that are
    zeroed at allocation only; a shorter batch in the same bucket must not see the previous batch's tail columns."""
    mine, ref = _branch_pair()
    B = 16
    g1, r1 = _branch_grads(mine, ref, B, 300, seed=1)   # M = 16 * 301 = 4816 -> NS = 4, ldt = 4864
    assert rel_err(g1, r1) < 3e-2
    g2, r2 = _branch_grads(mine, ref, B, 298, seed=2)   # M = 4784: same ldt, 32 stale columns unless they are cleared
    assert rel_err(g2, r2) < 3e-2, rel_err(g2, r2)
    g3, r3 = _branch_grads(mine, ref, B, 285, seed=3)   # M = 4576 -> ldt = 4608 (next bucket down), then back up
    assert rel_err(g3, r3) < 3e-2
    g4, r4 = _branch_grads(mine, ref, B, 287, seed=4)   # M = 4608 exactly fills its bucket
    assert rel_err(g4, r4) < 3e-2


def test_fused_adam_state_roundtrip_and_skipped_parameters():
    """FusedAdam's moments / step count travel through ``state_dict`` in torch.optim.Adam's layout (a
    resumed run continues bit for bit, torch-Adam state loads), and a parameter without gradient is left untouched."""
    from speechclip_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(33,), (17, 5), (4, 4, 3), ()]
    make = lambda: [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(i)).to(DEV)) for i, s in enumerate(shapes)]
    grads = [[torch.randn(s, generator=g).to(DEV) for s in shapes] for _ in range(5)]
    kw = dict(lr=1e-2, weight_decay=1e-2)

    def run(opt, params, steps, skip=None):
        for k in steps:
            for i, p in enumerate(params):
                p.grad = None if skip == i else grads[k][i].clone()
            opt.step()

    pa, pt = make(), make()
    a, t = FusedAdam(pa, **kw), torch.optim.Adam(pt, **kw)
    run(a, pa, range(3))
    run(t, pt, range(3))
    for x, y in zip(pa, pt):
        assert torch.allclose(x, y, atol=2e-6)
    # state_dict in torch's layout: FusedAdam -> torch Adam and torch Adam -> FusedAdam, then two more steps each
    sa, st = a.state_dict(), t.state_dict()
    assert set(sa["state"]) == set(st["state"]) == {0, 1, 2, 3}
    for i in range(4):
        assert int(sa["state"][i]["step"]) == 3 and torch.allclose(sa["state"][i]["exp_avg"], st["state"][i]["exp_avg"], atol=1e-6)
    pb, pu = make(), make()
    with torch.no_grad():
        for dst, src in zip(pb + pu, pt + pa):
            dst.copy_(src)
    b, u = FusedAdam(pb, **kw), torch.optim.Adam(pu, **kw)
    b.load_state_dict(st)
    u.load_state_dict(sa)
    run(a, pa, range(3, 5))
    run(b, pb, range(3, 5))
    run(u, pu, range(3, 5))
    run(t, pt, range(3, 5))
    for x, y, z, w in zip(pa, pb, pu, pt):
        assert torch.allclose(x, w, atol=3e-6) and torch.allclose(y, w, atol=3e-6) and torch.allclose(z, w, atol=3e-6)
    # a parameter whose grad is None: torch skips it (no decay, no moment update)
    before = pa[1].detach().clone()
    m_before = a.state_dict()["state"][1]["exp_avg"].clone()
    run(a, pa, [0], skip=1)
    run(t, pt, [0], skip=1)
    assert torch.equal(pa[1], before) and torch.equal(a.state_dict()["state"][1]["exp_avg"], m_before)
    for x, w in zip(pa, pt):
        assert torch.allclose(x, w, atol=4e-6)


def test_backward_after_a_second_graph_replay_is_refused():
    """under CUDA-graph replay the hidden-state slab the weighted sum saved for backward is the graph's
    output buffer; a second forward with the same signature overwrites it, so the first backward must fail loudly."""
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200 import engine
    from speechclip_b200.configs import parallel_config
    old = engine.GRAPHS
    engine.GRAPHS = True
    try:
        torch.manual_seed(0)
        model = KWClip_GeneralTransformer(OrderedNamespace(parallel_config("tiny"))).to(DEV).eval()
        g = torch.Generator().manual_seed(1)
        b = {"wav": (0.1 * torch.randn(4, 6000, generator=g)).to(DEV), "wav_len": torch.full((4,), 6000).to(DEV),
             "image": torch.randn(4, 3, 32, 32, generator=g).to(DEV), "id": torch.arange(4).to(DEV)}
        for _ in range(3):  # eager warm-up, capture, replay
            loss = model.training_step_end(model.training_step(b))["loss"]
            loss.backward()
        first = model.training_step_end(model.training_step(b))["loss"]
        second = model.training_step_end(model.training_step(b))["loss"]
        second.backward()  # latest forward: fine
        with pytest.raises(RuntimeError, match="overwrote the hidden states"):
            first.backward()
    finally:
        engine.GRAPHS = old
