"""The drop-in model (avssl.model.KWClip_GeneralTransformer on the CUDA path) against the CPU oracle, end to end:
forward features, masked InfoNCE, gradients of every trainable parameter, the fused clip+Adam step, retrieval indices,
and the inference entry points (encode_speech, feature_extractor_s3prl)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-8)).item()


def build(size, seed=0, warmup=5000):
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from speechclip_b200.configs import parallel_config
    cfg = parallel_config(size)
    cfg["audio_encoder"]["scheduler"]["warmup"] = warmup
    torch.manual_seed(seed)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    with torch.no_grad():  # non-trivial values for the parameters torch initialises to constants
        g = torch.Generator().manual_seed(seed + 1)
        model.audio_encoder.weightedsum_layer.weights.copy_(0.5 * torch.randn(model.audio_encoder.weightedsum_layer.weights.shape, generator=g))
        for n, p in model.parallel_branch.named_parameters():
            if n.endswith("bias") or "norm" in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    ta = cfg["model_settings"]["parallel_branch"]["transformer_args"]
    oracle = osc.SpeechClipOracle(oh.HubertCfg.named(cfg["audio_encoder"]["name"]), oc.ClipCfg.named(cfg["clip"]["name"]),
                                  dict(n_layers=1, nhead=ta["nhead"], dim_feedforward=ta["dim_feedforward"]),
                                  dict(temperature=0.07, temperature_trainable=cfg["cl_loss"]["args"]["temperature_trainable"]),
                                  normalize_hiddenstates=cfg["audio_encoder"]["normalize_hiddenstates"]).eval()
    missing, unexpected = oracle.load_state_dict(model.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return cfg, model.to(DEV), oracle


def batch(lens, size, seed=3, ids=None):
    g = torch.Generator().manual_seed(seed)
    wavs = [0.1 * torch.randn(n, generator=g) for n in lens]
    B = len(lens)
    img = torch.randn(B, 3, size, size, generator=g)
    ids = torch.arange(B) if ids is None else ids
    padded = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True)
    return wavs, img, ids, {"wav": padded.to(DEV), "wav_len": torch.tensor(lens).to(DEV), "image": img.to(DEV), "id": ids.to(DEV)}


@pytest.mark.parametrize("size", ["tiny", "tiny_large"])
def test_forward_loss_grads_match_oracle(size):
    cfg, model, oracle = build(size)
    model.eval()  # eval-mode arithmetic (no crop / dropout), gradients still flow to the trainable head
    lens = [6000, 4100, 5555, 6000, 3000, 4800, 6000, 5000]
    ids = torch.tensor([0, 1, 1, 2, 3, 3, 4, 5])  # same-id pairs exercise the negative mask
    wavs, img, ids, b = batch(lens, 32, ids=ids)
    out = model.training_step(b)
    feats = out["loss_feats"]
    loss = model.training_step_end(out)["loss"]
    loss.backward()

    of = oracle(wavs, img, ids)
    oloss, ologits = oracle.compute_loss(of, return_logits=True)
    oloss.backward()
    # embeddings: unit-norm rows; compare directions and values
    for k in ("parallel_audio_feat", "image_feat"):
        a, r = feats[k].detach().cpu(), of[k].detach()
        assert (a - r).abs().max() < 5e-3, (k, (a - r).abs().max())
    # logits within 1e-3 relative of the logit scale (north_star), loss within 1e-3 relative
    mult = oracle.criterion.multiplier()
    mult = float(mult.detach()) if torch.is_tensor(mult) else mult
    logits = feats["parallel_audio_feat"].detach().cpu() @ feats["image_feat"].detach().cpu().t() * mult
    assert (logits - ologits.detach()).abs().max() < 1e-3 * mult * 5, (logits - ologits).abs().max()
    assert abs(loss.item() - oloss.item()) < 2e-3 * max(1.0, abs(oloss.item())), (loss.item(), oloss.item())
    # gradients of every trainable parameter
    oparams = dict(oracle.named_parameters())
    checked = 0
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        og = oparams[name].grad
        assert p.grad is not None and og is not None, name
        e = rel_err(p.grad.cpu(), og)
        assert e < 3e-2, (name, e)
        checked += 1
    assert checked == (19 if size == "tiny_large" else 18), checked


def test_retrieval_indices_bit_exact_vs_oracle_on_same_embeddings():
    """argmax retrieval on the CUDA path == the oracle's on the same embeddings (north_star: bit-exact indices)."""
    cfg, model, oracle = build("tiny")
    model.eval()
    lens = [6000] * 16
    wavs, img, ids, b = batch(lens, 32, seed=5)
    with torch.no_grad():
        _, _, others = model(b)
    a, i = others["parallel_audio_feat"], others["image_feat"]
    from speechclip_b200 import ops
    score = torch.empty(16, 16, device=DEV)
    ops.sgemm(a.contiguous(), i.contiguous(), score)
    top1 = torch.empty(16, device=DEV, dtype=torch.int32)
    ops.retrieval_rank(score, None, None, None, top1)
    ref = (a.cpu() @ i.cpu().t()).argmax(1)
    assert torch.equal(top1.cpu().long(), ref)
    out = model.validation_epoch_end([{"id": ids, "audio_feat": a, "image_feat": i}])
    from oracle import speechclip as osc
    s = a.cpu() @ i.cpu().t()
    ref_r = osc.mutual_retrieval(s, s.t().contiguous(), ids, ids, [1, 5, 10])
    for m_, r_ in zip(out, ref_r):
        for k in r_:
            assert abs(m_[k] - r_[k]) < 1e-4


def test_training_steps_follow_torch_adam_on_the_oracle():
    """3 optimizer steps (global-norm clip 4.0 + Adam + linear warm-up schedule) track the oracle trained with torch."""
    cfg, model, oracle = build("tiny", warmup=2)
    model.eval()
    opts, scheds = model.configure_optimizers()
    opt, sched = opts[0], scheds[0]["scheduler"]
    start = {n: p.detach().cpu().clone() for n, p in model.named_parameters() if p.requires_grad}
    oparams = [p for n, p in oracle.named_parameters() if n.startswith("parallel_branch") or "weightedsum" in n]
    oopt = torch.optim.Adam(oparams, lr=1e-4, weight_decay=1e-6)
    from oracle import speechclip as osc
    losses, olosses = [], []
    for step in range(3):
        wavs, img, ids, b = batch([5000, 6000, 4000, 6000], 32, seed=10 + step)
        out = model.training_step(b)
        loss = model.training_step_end(out)["loss"]
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        losses.append(loss.item())
        ol = oracle.compute_loss(oracle(wavs, img, ids))
        oopt.zero_grad()
        ol.backward()
        torch.nn.utils.clip_grad_norm_(oparams, 4.0)
        for gparam in oopt.param_groups:
            gparam["lr"] = 1e-4 * osc.linear_warmup_decay(step, 1e-4, 2, 50000, 1e-8)
        oopt.step()
        olosses.append(ol.item())
    for a, r in zip(losses, olosses):
        assert abs(a - r) < 3e-3 * max(1.0, abs(r)), (losses, olosses)
    od = dict(oracle.named_parameters())
    for n, p in model.named_parameters():
        if p.requires_grad:
            mine, ref = p.detach().cpu() - start[n], od[n].detach() - start[n]
            if n.endswith("in_proj_bias"):
                # d loss / d b_k is exactly zero (a key bias shifts every score of a row equally): both sides apply
                # sign(noise) * lr there, so leave the K third out of the comparison
                d3 = mine.numel() // 3
                keep = torch.cat([torch.arange(0, d3), torch.arange(2 * d3, 3 * d3)])
                mine, ref = mine[keep], ref[keep]
            # Adam's early steps are sign-like (|update| ~ lr per element): elements whose gradient is at noise level may
            # differ, so compare the update in the mean, relative to the mean update size
            assert ref.abs().mean() > 1e-5, n
            assert (mine - ref).abs().mean() < 0.1 * ref.abs().mean(), (n, (mine - ref).abs().mean().item(), ref.abs().mean().item())
    assert model.arena().intact()


def test_inference_entry_points():
    cfg, model, oracle = build("tiny")
    model.eval()
    g = torch.Generator().manual_seed(9)
    wavs = [0.1 * torch.randn(n, generator=g) for n in (6000, 3500)]
    with torch.no_grad():
        enc = model.encode_speech([w.to(DEV) for w in wavs])
        last, hidden = model.feature_extractor_s3prl([w.to(DEV) for w in wavs])
        oe = oracle.encode_speech(wavs)
        olast, ohidden = oracle.feature_extractor_s3prl(wavs)
    assert (enc["parallel_audio_feat"].cpu() - oe["parallel_audio_feat"]).abs().max() < 5e-3
    assert len(hidden) == len(ohidden) == 4  # 3 HuBERT states + 1 branch output (example.py:29 -> 14 for base)
    valid = torch.tensor([[True] * 18, [True] * 11 + [False] * 7])
    for a, r in zip(hidden, ohidden):
        assert a.shape == r.shape
        assert rel_err(a.cpu()[valid], r[valid]) < 1.5e-2
    assert torch.equal(last, hidden[-1])


def test_cpu_tensors_raise():
    cfg, model, oracle = build("tiny")
    with pytest.raises(RuntimeError):
        model.cpu().forward({"wav": torch.randn(1, 4000), "wav_len": torch.tensor([4000]), "image": torch.randn(1, 3, 32, 32),
                             "id": torch.tensor([0])})
