"""End-to-end parity at the benchmark's own geometry: ``KWClip_GeneralTransformer`` on the CUDA path
against the CPU oracle with FULL-SIZE towers and 102 400-sample utterances (T = 319 frames, spchclp_p.yaml:104), so that the
CTA-pair GEMMs, the TMA-store epilogues, the tcgen05 attention at 319 keys and the positional-conv slab path are all live
together: per-layer hidden states, weighted sum, pooled embeddings, logits, loss, every gradient, and top-1 retrieval indices of
the CUDA path against the ORACLE path (reference: kwClip.py:1385-1478, losses.py:185-245, retrieval.py:45-65).

Tolerances (north_star: logits <= 1e-3 relative, retrieval indices bit-exact):
* logits: |logit - oracle| <= 1e-3 x the logit scale (= 1/temperature: logits are cosines x scale, so this bounds the cosine
  error by 1e-3);
* retrieval: top-1 indices equal the oracle's, except rows where the oracle's own best and second-best scores are closer than
  twice the measured cosine error (a provable tie); the number of such rows is reported and bounded;
* hidden states: max-abs error relative to the state's max-abs value (fp16 operands / fp16 hidden stream against fp32).
Every measured error is written to gpurun_out/parity_fullsize_<case>.json.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL = dict(hidden=1.5e-2, audio_feat=1.0e-2, embedding=1.0e-3, logits_rel=1.0e-3, loss_rel=1.0e-3, grad=3.0e-2)
N_SAMPLES = 102400


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-8)).item()


def _report(case, rep):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"parity_fullsize_{case}.json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(f"[parity {case}] " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in rep.items() if not isinstance(v, dict)))


def _build_parallel(size, seed=0):
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from speechclip_b200.configs import parallel_config
    cfg = parallel_config(size)
    torch.manual_seed(seed)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    with torch.no_grad():
        g = torch.Generator().manual_seed(seed + 1)
        ws = model.audio_encoder.weightedsum_layer.weights
        ws.copy_(0.5 * torch.randn(ws.shape, generator=g))
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
    return cfg, model.to(DEV).eval(), oracle


def _batch(lens, ids, seed):
    g = torch.Generator().manual_seed(seed)
    wavs = [0.1 * torch.randn(n, generator=g) for n in lens]
    img = torch.randn(len(lens), 3, 224, 224, generator=g)
    ids = torch.tensor(ids)
    padded = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True)
    return wavs, img, ids, {"wav": padded.to(DEV), "wav_len": torch.tensor(lens).to(DEV), "image": img.to(DEV), "id": ids.to(DEV)}


def _top1_vs_oracle(mine_a, mine_i, ref_a, ref_i, cos_err, rep, key):
    """Top-1 retrieval of the CUDA path (its own embeddings through scb_sgemm + scb_retrieval_rank) against the oracle path
    (the oracle's embeddings, torch argmax).  A differing row must be a provable tie of the oracle's own scores."""
    from speechclip_b200 import ops
    n = mine_a.shape[0]
    ties = 0
    for tag, (qa, ca, qr, cr) in {"A2I": (mine_a, mine_i, ref_a, ref_i), "I2A": (mine_i, mine_a, ref_i, ref_a)}.items():
        score = torch.empty(n, n, device=DEV)
        ops.sgemm(qa.contiguous(), ca.contiguous(), score)
        top1 = torch.empty(n, device=DEV, dtype=torch.int32)
        ops.retrieval_rank(score, None, None, None, top1)
        ref_score = qr @ cr.t()
        ref_top = ref_score.argmax(1)
        diff = (top1.cpu().long() != ref_top).nonzero().view(-1).tolist()
        for r in diff:
            gap = (ref_score[r, ref_top[r]] - ref_score[r, top1[r].item()]).item()
            assert 0 <= gap <= 2 * cos_err, (tag, r, gap, cos_err)
        ties += len(diff)
    rep[key] = ties
    return ties


CASES = {
    # Parallel-base at T = 319 with Flickr8k-style same-id groups (5 captions per image -> masked negatives)
    "base_full": dict(size="base", lens=[N_SAMPLES] * 10, ids=[0, 0, 0, 1, 1, 2, 3, 3, 4, 5]),
    # ragged lengths: padding masks, round(len / 320), zeroed frames before the positional conv
    "base_ragged": dict(size="base", lens=[N_SAMPLES, 33000, 64321, 80000, 47999, N_SAMPLES, 96000, 51200], ids=list(range(8))),
    # Parallel-large: LayerNorm extractor, pre-LN, normalised waveform / hidden states, learnable temperature, ViT-L/14
    "large": dict(size="large", lens=[N_SAMPLES, 70000, N_SAMPLES, 88888], ids=[0, 1, 1, 2]),
}


@pytest.mark.parametrize("case", list(CASES))
def test_parallel_full_size_forward_loss_grads_retrieval(case):
    c = CASES[case]
    cfg, model, oracle = _build_parallel(c["size"])
    wavs, img, ids, b = _batch(c["lens"], c["ids"], seed=21)
    B = len(wavs)
    rep = {"B": B, "size": c["size"]}

    # ---- tower: every hidden state + the weighted sum
    with torch.no_grad():
        feat, feat_len, states = model.forward_audio(b["wav"], b["wav_len"], return_hidden_states=True)
        ofeat, olen, ostates = oracle.forward_audio(wavs, return_hidden_states=True)
    assert feat.shape == ofeat.shape and feat.shape[1] == 319 and feat_len.cpu().tolist() == olen.tolist()
    valid = torch.arange(319)[None, :] < olen[:, None]
    errs = [rel_err(a.cpu()[valid], r[valid]) for a, r in zip(states, ostates)]
    rep["hidden_rel_max"], rep["hidden_rel_last"] = max(errs), errs[-1]
    rep["audio_feat_rel"] = rel_err(feat.cpu()[valid], ofeat[valid])
    assert len(states) == len(ostates) == (25 if c["size"] == "large" else 13)

    # ---- the training step
    out = model.training_step(b)
    feats = out["loss_feats"]
    loss = model.training_step_end(out)["loss"]
    loss.backward()
    of = oracle(wavs, img, ids)
    oloss, ologits = oracle.compute_loss(of, return_logits=True)
    oloss.backward()
    ma, mi = feats["parallel_audio_feat"].detach(), feats["image_feat"].detach()
    ra, ri = of["parallel_audio_feat"].detach(), of["image_feat"].detach()
    rep["audio_emb_abs"] = (ma.cpu() - ra).abs().max().item()
    rep["image_emb_abs"] = (mi.cpu() - ri).abs().max().item()
    mult = oracle.criterion.multiplier()
    mult = float(mult.detach()) if torch.is_tensor(mult) else float(mult)
    logits = torch.empty(B, B, device=DEV)
    from speechclip_b200 import ops
    ops.sgemm(ma.contiguous(), mi.contiguous(), logits)
    rep["logits_rel"] = ((logits.cpu() * mult - ologits.detach()).abs().max() / mult).item()   # = max cosine error
    rep["loss"], rep["oracle_loss"] = loss.item(), oloss.item()
    rep["loss_rel"] = abs(loss.item() - oloss.item()) / max(1.0, abs(oloss.item()))
    gerrs = {}
    oparams = dict(oracle.named_parameters())
    for name, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and oparams[name].grad is not None, name
            gerrs[name] = rel_err(p.grad.cpu(), oparams[name].grad)
    rep["grad_rel_max"] = max(gerrs.values())
    rep["grad_worst"] = max(gerrs, key=gerrs.get)
    rep["n_grads"] = len(gerrs)
    _top1_vs_oracle(ma, mi, ra, ri, rep["logits_rel"], rep, "retrieval_tie_rows")
    _report(case, rep)

    assert rep["hidden_rel_max"] < TOL["hidden"], rep
    assert rep["audio_feat_rel"] < TOL["audio_feat"], rep
    assert rep["audio_emb_abs"] < TOL["embedding"] and rep["image_emb_abs"] < TOL["embedding"], rep
    assert rep["logits_rel"] < TOL["logits_rel"], rep
    assert rep["loss_rel"] < TOL["loss_rel"], rep
    assert rep["grad_rel_max"] < TOL["grad"], rep
    assert rep["n_grads"] == (19 if c["size"] == "large" else 18)
    assert rep["retrieval_tie_rows"] <= 1, rep


def test_cascaded_base_full_size_vs_oracle():
    """Cascaded SpeechCLIP-base at T = 319 with the 8112-entry reduced vocabulary: the selected keyword ids equal the oracle's
    except provable ties (counted), then — with the oracle held to the CUDA path's ids — features, loss, every gradient, and the
    top-1 retrieval indices against the oracle path."""
    import tempfile
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import cascaded_config, write_synthetic_vocab_usage
    K = 8
    with tempfile.TemporaryDirectory() as td:
        cfg = cascaded_config("base", write_synthetic_vocab_usage(os.path.join(td, "usage.npy")))
        cfg["model_settings"]["cascaded_branch"]["transformer_args"]["dropout"] = 0.0   # dropout has its own test (test_dropout_gpu.py)
        torch.manual_seed(0)
        model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        ws = model.audio_encoder.weightedsum_layer.weights
        ws.copy_(0.5 * torch.randn(ws.shape, generator=g))
    sot, eot = model.clip.special_tokens()
    ccfg = oc.ClipCfg.named("ViT-B/32")
    ccfg.vocab = model.clip.model.token_embedding.weight.shape[0]
    assert ccfg.vocab == 8112
    oracle = osc.SpeechClipOracle(oh.HubertCfg.named("hubert"), ccfg, None, dict(temperature=0.07, temperature_trainable=False),
                                  cascaded_args=dict(keyword_num=K, nhead=1, vq_temp=0.1, sot_token=sot, eot_token=eot)).eval()
    sd = {k: v for k, v in model.state_dict().items() if not k.startswith("cascaded_branch.clip.") and "vector_quantizer" not in k
          and k != "clip.original_text_emb_weight"}
    missing, unexpected = oracle.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    oracle.cascaded_branch.bn_layer.train()
    model = model.to(DEV).train()   # train mode: BatchNorm batch statistics + straight-through softmax (102400 samples: no crop)

    B = 8
    wavs, img, ids, b = _batch([N_SAMPLES] * 5 + [90000, 64000, 77777], [0, 0, 1, 2, 3, 4, 4, 5], seed=31)
    rep = {"B": B, "size": "cascaded-base", "vocab": 8112}
    feats, log_metrics, others = model(b)
    loss = model.training_step_end({"loss_feats": feats, "log_metrics": log_metrics})["loss"]
    loss.backward()
    with torch.no_grad():
        free = oracle(wavs, img, ids)   # the oracle on its own: which ids does an fp32 evaluation select?
    mine = others["vq_results"]["targets"].cpu().view(-1)
    ref_ids = free["vq_results"]["targets"].view(-1)
    ref_scores = free["cascaded_collect"]["cos"].detach().view(mine.numel(), -1).clone()
    ref_scores[:, [0, 2, 3]] = float("-inf")
    diff = (mine != ref_ids).nonzero().view(-1).tolist()
    gaps = [(ref_scores[r, ref_ids[r]] - ref_scores[r, mine[r]]).item() for r in diff]
    top2 = ref_scores.topk(2, dim=1).values
    rep["vq_rows"], rep["vq_id_mismatch_rows"], rep["vq_mismatch_gap_max"] = mine.numel(), len(diff), max(gaps) if gaps else 0.0
    rep["oracle_top2_gap_min"] = (top2[:, 0] - top2[:, 1]).min().item()
    oracle.force_idx = mine
    of = oracle(wavs, img, ids)
    oloss = oracle.compute_loss(of)
    oloss.backward()
    ma, mi = feats["cascaded_audio_feat"].detach(), feats["image_feat"].detach()
    ra, ri = of["cascaded_audio_feat"].detach(), of["image_feat"].detach()
    rep["audio_emb_abs"] = (ma.cpu() - ra).abs().max().item()
    rep["image_emb_abs"] = (mi.cpu() - ri).abs().max().item()
    rep["logits_rel"] = (ma.cpu() @ mi.cpu().t() - ra @ ri.t()).abs().max().item()
    rep["loss"], rep["oracle_loss"] = loss.item(), oloss.item()
    rep["loss_rel"] = abs(loss.item() - oloss.item()) / max(1.0, abs(oloss.item()))
    oparams = dict(oracle.named_parameters())
    # Gradients.  Train-mode BatchNorm removes, to first order, anything that shifts a keyword feature by a constant over the batch:
    # the TRUE gradients of the parameters in front of it that act that way ([CLS] rows, attention / projection biases, the
    # LayerNorm bias) are two to five orders of magnitude below the others and sit at the rounding floor of the bf16 / TF32
    # backward GEMMs (a flat ~1.3e-6 absolute here).  Every gradient is therefore bounded relative to max(its own magnitude,
    # 2 % of the largest gradient magnitude of the model); the raw numbers are written to the report.
    gerrs, detail = {}, {}
    gmax = max(oparams[name].grad.abs().max().item() for name, p in model.named_parameters() if p.requires_grad)
    for name, p in model.named_parameters():
        if p.requires_grad:
            og = oparams[name].grad
            err = (p.grad.cpu() - og).abs().max().item()
            detail[name] = [err, og.abs().max().item(), p.grad.abs().max().item()]
            gerrs[name] = err / max(og.abs().max().item(), 0.02 * gmax)
    rep["grad_rel_max"], rep["grad_worst"], rep["n_grads"] = max(gerrs.values()), max(gerrs, key=gerrs.get), len(gerrs)
    rep["grad_detail_abs_err__ref_max__mine_max"] = detail
    _top1_vs_oracle(ma, mi, ra, ri, rep["logits_rel"], rep, "retrieval_tie_rows")
    _report("cascaded_base", rep)
    # index work: equal ids except provable ties (the oracle's own top-2 gap below the score error of the fp16 towers)
    for gap in gaps:
        assert 0 <= gap < 2e-3, (gaps, diff)
    assert len(diff) <= 2, rep
    assert rep["audio_emb_abs"] < 5e-3 and rep["image_emb_abs"] < TOL["embedding"], rep
    assert rep["loss_rel"] < 2e-3, rep
    assert rep["grad_rel_max"] < 4e-2 and rep["n_grads"] == 12, rep
    assert rep["retrieval_tie_rows"] <= 1, rep


@pytest.mark.parametrize("hidden", ["fp16", "fp32"])
def test_hubert_base_heavy_tailed_channels_in_the_hidden_stream(hidden, monkeypatch):
    """Trained HuBERT checkpoints carry outlier channels (a few LayerNorm gains / biases far above the rest).  The post-LN tower
    keeps its hidden states in fp16 only by default (engine.py HubertPlan._forward; ``SCB_HIDDEN_FP32=1`` keeps an fp32 stream):
    scale a few gains x50 and biases +-30 in several layers and compare with the fp32 oracle.  The yardstick is the reference's
    OWN arithmetic: it trains and evaluates under fp16 autocast (trainer.precision: 16, spchclp_p.yaml:113) — the same oracle under
    torch's CPU autocast(float16) (fp16 GEMM inputs and outputs, fp32 LayerNorm / softmax / residual stream).  Required: within 1.5x
    of that error on the worst state and below it per channel, in BOTH modes.  Measured (gpurun_out/parity_fullsize_heavy_tail_*.json):
    the fp16 and fp32 hidden streams give the same error (2.12e-2 vs 2.13e-2 of the state maximum, 0.16 vs 0.14 per channel; the
    reference's autocast: 1.71e-2, 0.24) — what limits accuracy with outlier channels is the fp16 rounding of the GEMM operands,
    which the reference shares, not the precision of the hidden-state stream."""
    from avssl.module import FairseqSpeechEncoder_Hubert
    from oracle import hubert as oh
    from speechclip_b200 import engine
    from speechclip_b200.init import seeded_init_
    monkeypatch.setattr(engine, "HIDDEN16", hidden == "fp16")
    enc = FairseqSpeechEncoder_Hubert("hubert", feat_select_idx="hidden_states")
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for l in (0, 3, 7, 11):
            for ln in ("self_attn_layer_norm", "final_layer_norm"):
                mod = getattr(enc.encoder.encoder.layers, str(l))
                w, bias = getattr(mod, ln).weight, getattr(mod, ln).bias
                idx = torch.randperm(768, generator=g)[:6]
                w[idx] *= 50.0
                bias[idx] += 30.0 * torch.sign(torch.randn(6, generator=g))
        idx = torch.randperm(768, generator=g)[:6]
        enc.encoder.encoder.layer_norm.weight[idx] *= 50.0
    om = seeded_init_(oh.HubertModel(oh.HubertCfg.named("hubert")), 7122).eval()
    om.load_state_dict(enc.encoder.state_dict())
    enc = enc.to(DEV).eval()
    wav = 0.1 * torch.randn(4, 48000, generator=torch.Generator().manual_seed(6))
    states, _ = enc(wav.to(DEV))
    with torch.no_grad():
        ref = om.custom_forward(wav, None)["layer_results"]
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.float16):
        amp = [a.float() for a in om.custom_forward(wav, None)["layer_results"]]

    def errors(states):
        worst, worst_ch = 0.0, 0.0
        for a, r in zip(states, ref):
            worst = max(worst, rel_err(a, r))
            ch_scale = r.abs().amax(dim=(0, 1)).clamp_min(1e-3)            # per-channel magnitude: outliers must not hide the rest
            worst_ch = max(worst_ch, ((a - r).abs().amax(dim=(0, 1)) / ch_scale).max().item())
        return worst, worst_ch

    mine = [a.cpu() for a in states]
    assert all(torch.isfinite(a).all() for a in mine)
    worst, worst_ch = errors(mine)
    amp_worst, amp_worst_ch = errors(amp)
    rep = {"hidden_stream": hidden, "hidden_rel_max": worst, "hidden_rel_per_channel_max": worst_ch, "autocast_fp16_hidden_rel_max": amp_worst,
           "autocast_fp16_hidden_rel_per_channel_max": amp_worst_ch, "state_abs_max": max(r.abs().max().item() for r in ref)}
    _report("heavy_tail_" + hidden, rep)
    assert rep["state_abs_max"] > 100.0   # the outliers are really there
    assert worst <= max(1.5e-2, 1.5 * amp_worst) and worst_ch <= max(4e-2, amp_worst_ch), rep
