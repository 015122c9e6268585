"""Cascaded branch on the CUDA path against the CPU oracle: the new kernels one by one, the CLIP text
tower (pruned to the K+2 live positions) forward and backward, KW_CascadedBranch end to end, and the cascaded model's
training step.  Index work (the selected vocabulary ids) must be bit-exact; floating point within the stated tolerances."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-8)).item()


def _assert_same_ids(mine, ref, ref_scores, tie=2e-5, min_agree=1.0):
    """Selected ids must equal the oracle's; a difference is only tolerated where the oracle's own scores of the two
    candidates are within ``tie`` of each other (a rounding-level tie), and at most 1 - min_agree of the rows."""
    mine, ref = mine.view(-1), ref.view(-1)
    diff = (mine != ref).nonzero().view(-1)
    for r in diff.tolist():
        gap = (ref_scores[r, ref[r]] - ref_scores[r, mine[r]]).item()
        assert 0 <= gap < tie, (r, int(mine[r]), int(ref[r]), gap)
    assert len(diff) <= (1.0 - min_agree) * len(ref), (len(diff), len(ref))
    return len(diff)


# ---------------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("heads,hd,nq,Tk,B", [(1, 768, 8, 57, 5), (1, 1024, 8, 335, 3), (4, 64, 3, 40, 4), (2, 96, 1, 9, 2)])
def test_mq_attention_fwd_bwd(heads, hd, nq, Tk, B):
    from speechclip_b200 import ops
    g = torch.Generator().manual_seed(1)
    d = heads * hd
    q = torch.randn(nq, d, generator=g)
    kv = (0.5 * torch.randn(B, Tk, 2 * d, generator=g)).half()
    lens = torch.randint(max(1, Tk // 2), Tk + 1, (B,), generator=g)
    lens[0] = Tk
    dctx = torch.randn(B, nq, d, generator=g)
    # fp32 reference on the fp16-rounded K/V
    qr = q.clone().requires_grad_(True)
    kvr = kv.float().requires_grad_(True)
    k_, v_ = kvr[..., :d].view(B, Tk, heads, hd), kvr[..., d:].view(B, Tk, heads, hd)
    s = torch.einsum("qhd,bthd->bhqt", qr.view(nq, heads, hd), k_) * hd ** -0.5
    mask = torch.arange(Tk)[None, :] >= lens[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    ctx = torch.einsum("bhqt,bthd->bqhd", p, v_).reshape(B, nq, d)
    ctx.backward(dctx)

    qd, kvd = q.to(DEV), kv.to(DEV)
    kv_len = lens.to(DEV, torch.int32)
    probs = torch.empty(B, heads, nq, Tk, device=DEV)
    out = torch.empty(B, nq, d, device=DEV)
    ops.mq_attention_fwd(qd, kvd, 0, d, kv_len, heads, hd, hd ** -0.5, probs, out)
    assert rel_err(out.cpu(), ctx.detach()) < 2e-5
    assert (probs.cpu() - p.detach()).abs().max() < 2e-6
    dkv = torch.full((B, Tk, 2 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
    dq_part = torch.full((B, nq * d), float("nan"), device=DEV)
    ops.mq_attention_bwd(qd, kvd, 0, d, kv_len, heads, hd, hd ** -0.5, probs, dctx.to(DEV), dkv, dq_part)
    assert rel_err(dq_part.sum(0).view(nq, d).cpu(), qr.grad) < 1e-4
    assert rel_err(dkv.float().cpu(), kvr.grad) < 6e-3   # bf16 output rounding
    assert torch.isfinite(dkv.float()).all()


@pytest.mark.parametrize("B,K,W", [(16, 8, 512), (5, 3, 40), (1, 8, 64)])
def test_keyword_batchnorm_matches_torch(B, K, W):
    from speechclip_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, K, W, generator=g) * 2 + 0.3
    bn = torch.nn.BatchNorm1d(K * W)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(K * W, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(K * W, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(K * W, generator=g))
        bn.running_var.copy_(torch.rand(K * W, generator=g) + 0.5)
    rm, rv = bn.running_mean.clone().to(DEV), bn.running_var.clone().to(DEV)
    xr = x.clone().requires_grad_(True)
    if B > 1:
        y = bn(xr.permute(0, 2, 1).reshape(B, -1)).reshape(B, W, K).permute(0, 2, 1)
        dy = torch.randn(B, K, W, generator=g)
        y.backward(dy)
        xd, yd = x.to(DEV), torch.empty(B, K, W, device=DEV)
        mean, rstd = torch.empty(K * W, device=DEV), torch.empty(K * W, device=DEV)
        wd, bd = bn.weight.detach().to(DEV), bn.bias.detach().to(DEV)
        ops.batchnorm_fwd(xd, yd, wd, bd, rm, rv, mean, rstd, 1e-5, 0.1, True)
        assert rel_err(yd.cpu(), y.detach()) < 1e-5
        assert rel_err(rm.cpu(), bn.running_mean) < 1e-5 and rel_err(rv.cpu(), bn.running_var) < 1e-5
        dx, dg, db = torch.empty_like(xd), torch.empty(K * W, device=DEV), torch.empty(K * W, device=DEV)
        ops.batchnorm_bwd(dy.to(DEV), xd, wd, mean, rstd, dx, dg, db)
        assert rel_err(dx.cpu(), xr.grad) < 1e-4
        assert rel_err(dg.cpu(), bn.weight.grad) < 1e-5 and rel_err(db.cpu(), bn.bias.grad) < 1e-5
    bn.eval()
    ye = bn(x.permute(0, 2, 1).reshape(B, -1)).reshape(B, W, K).permute(0, 2, 1)
    yd = torch.empty(B, K, W, device=DEV)
    ops.batchnorm_fwd(x.to(DEV), yd, bn.weight.detach().to(DEV), bn.bias.detach().to(DEV), bn.running_mean.to(DEV), bn.running_var.to(DEV),
                      None, None, 1e-5, 0.1, False)
    assert rel_err(yd.cpu(), ye.detach()) < 1e-5


def test_split_tf32_gemm_is_fp32_accurate():
    from speechclip_b200 import ops
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(300, 512, generator=g), torch.randn(808, 512, generator=g)
    ref = (a.double() @ b.double().t())
    a3 = ops.split_tf32(a.to(DEV), torch.empty(300, 1536, device=DEV), 0)
    b3 = ops.split_tf32(b.to(DEV), torch.empty(808, 1536, device=DEV), 1)
    out = ops.gemm(a3, b3, out=torch.empty(300, 808, device=DEV))
    plain = ops.gemm(a.to(DEV), b.to(DEV), out=torch.empty(300, 808, device=DEV))
    e3 = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    e1 = (plain.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    # fp32-class: what is left is the tensor core's own fp32 accumulation over k = 1536 (measured 7e-6 of the largest entry)
    assert e3 < 2e-5, e3
    assert e1 > 10 * e3, (e1, e3)   # the plain TF32 product is what the split is there to beat


@pytest.mark.parametrize("R,V,W,temp", [(64, 8112, 512, 0.1), (24, 96, 64, 0.1), (7, 1000, 768, 0.5)])
def test_cosine_vq_forward_backward_match_oracle(R, V, W, temp):
    """Selected ids bit-exact; cosine scores, straight-through gradient and diagnostics within fp32 tolerances."""
    from oracle import speechclip as osc
    from speechclip_b200 import ops
    from speechclip_b200.cascaded import Vocabulary
    from speechclip_b200.engine import Workspace
    from speechclip_b200.cascaded import CascadedHead
    g = torch.Generator().manual_seed(4)
    E = 0.02 * torch.randn(V, W, generator=g) + 0.01
    kw = (0.03 * torch.randn(1, R, W, generator=g) + 0.01).requires_grad_(True)
    cos = torch.stack([F.cosine_similarity(kw[:, i, :].unsqueeze(-1), E.t().unsqueeze(0), dim=1) for i in range(R)], 1)
    vq = osc.simple_vector_quantizer(cos, temp, True)
    keywords = vq["subword_prob"] @ E
    dkeys = torch.randn(1, R, W, generator=g)
    keywords.backward(dkeys)

    vocab = Vocabulary(E, DEV)
    head = CascadedHead(64, 1, R, W)
    ws = Workspace(DEV)
    kwd = kw.detach().to(DEV).contiguous()
    cosd, idx, stats = head.quantize(ws, vocab, kwd, temp)
    ref_cos = cos.detach().view(R, V).clone()
    ref_cos[:, [0, 2, 3]] = float("-inf")
    got = cosd.cpu()
    assert torch.equal(torch.isinf(got), torch.isinf(ref_cos))
    fin = ~torch.isinf(ref_cos)
    assert (got[fin] - ref_cos[fin]).abs().max() < 1e-5
    _assert_same_ids(idx.view(-1).cpu(), vq["targets"].view(-1), ref_cos)
    # backward
    gbuf = torch.empty(R, (V + 3) // 4 * 4, device=DEV)[:, :V]
    ops.gemm(dkeys.view(R, W).to(DEV).contiguous(), vocab.E, out=gbuf)
    t2 = torch.empty(R, device=DEV)
    ops.vq_backward(gbuf, cosd, stats, temp, t2)
    t1 = torch.empty(R, W, device=DEV)
    ops.gemm(gbuf, vocab.unit_t[:, :V], out=t1)
    dkw = torch.empty(R, W, device=DEV)
    ops.cosine_bwd_rows(t1, t2, kwd.view(R, W), stats, dkw)
    assert rel_err(dkw.cpu(), kw.grad.view(R, W)) < 5e-3    # TF32 gradient GEMMs
    # diagnostics
    hist, avg, ent = torch.zeros(V, device=DEV), torch.zeros(V, device=DEV), torch.empty(R, device=DEV)
    ops.vq_diagnostics(cosd, stats, idx, hist, avg, ent)
    hp, ap = hist.cpu() / R, avg.cpu() / R
    assert abs(torch.exp(-(hp * torch.log(hp + 1e-7)).sum()).item() - vq["code_perplexity"].item()) < 1e-3 * vq["code_perplexity"].item()
    assert abs(torch.exp(-(ap * torch.log(ap + 1e-7)).sum()).item() - vq["prob_perplexity"].item()) < 1e-3 * vq["prob_perplexity"].item()
    assert abs(ent.mean().item() - vq["ent_per_t"].mean().item()) < 1e-4 * vq["ent_per_t"].mean().item()


def test_standalone_vector_quantizer_module():
    from avssl.module.speechclip_c_modules.my_vector_quantizer import SimpleVectorQuantizer
    from oracle import speechclip as osc
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 8, 200, generator=g)
    w = torch.randn(6, 8, 200, generator=g)
    for training in (True, False):
        xr = x.clone().requires_grad_(True)
        ref = osc.simple_vector_quantizer(xr, 0.1, training)
        vq = SimpleVectorQuantizer(temp="fixed=0.1").to(DEV).train(training)
        xd = x.clone().to(DEV).requires_grad_(True)
        out = vq(xd)
        assert torch.equal(out["targets"].cpu(), ref["targets"])
        # the oracle's hard + soft - soft.detach() leaves rounding residue around the one-hot values in training mode
        assert (out["subword_prob"].detach().cpu() - ref["subword_prob"].detach()).abs().max() < (1e-6 if training else 1e-12)
        for k in ("code_perplexity", "prob_perplexity", "diversity_loss"):
            assert abs(out[k].item() - ref[k].item()) < 1e-3 * max(1.0, abs(ref[k].item())), k
        assert (out["ent_per_t"].cpu() - ref["ent_per_t"]).abs().max() < 1e-4
        assert out["temp"] == pytest.approx(0.1) and out["num_vars"] == 200
        if training:
            (ref["subword_prob"] * w).sum().backward()
            (out["subword_prob"] * w.to(DEV)).sum().backward()
            assert rel_err(xd.grad.cpu(), xr.grad) < 1e-4


@pytest.mark.parametrize("B,L,heads,hd,causal", [(3, 10, 8, 64, True), (2, 10, 4, 16, True), (2, 7, 2, 32, False), (1, 33, 2, 64, True)])
def test_attention_small_bwd(B, L, heads, hd, causal):
    from speechclip_b200 import ops
    g = torch.Generator().manual_seed(6)
    d = heads * hd
    qkv = (0.7 * torch.randn(B, L, 3 * d, generator=g)).half()
    dctx = torch.randn(B, L, d, generator=g)
    r = qkv.float().requires_grad_(True)
    q, k, v = (r[..., i * d:(i + 1) * d].view(B, L, heads, hd).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) * hd ** -0.5
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool), 1), float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, d)
    o.backward(dctx)
    out = torch.empty(B, L, 3 * d, device=DEV)
    ops.attention_small_bwd(qkv.to(DEV), dctx.to(DEV), out, B, L, heads, hd, hd ** -0.5, causal)
    assert rel_err(out.cpu(), r.grad) < 1e-4
    # and the forward kernel on the same short causal shape
    ctx = torch.empty(B, L, d, device=DEV, dtype=torch.float16)
    qd = qkv.to(DEV)
    ops.attention(qd[:, :, :d], qd[:, :, d:2 * d], qd[:, :, 2 * d:], ctx, heads, hd ** -0.5, None, causal)
    assert rel_err(ctx.float().cpu(), o.detach()) < 3e-3


def test_activation_kernels():
    from speechclip_b200 import ops
    g = torch.Generator().manual_seed(7)
    pre = (2 * torch.randn(37, 64, generator=g)).half()
    dy = torch.randn(37, 64, generator=g)
    for act, fn in ((ops.ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x)), (ops.ACT_GELU, F.gelu)):
        r = pre.float().requires_grad_(True)
        y = fn(r)
        y.backward(dy)
        out = torch.empty(37, 64, device=DEV, dtype=torch.float16)
        ops.act16_fwd(pre.to(DEV), act, out)
        assert (out.float().cpu() - y.detach()).abs().max() < 2e-3
        dx = torch.empty(37, 64, device=DEV)
        ops.act_bwd(dy.to(DEV), pre.to(DEV), act, dx)
        assert rel_err(dx.cpu(), r.grad) < 1e-5


def test_softmax_rows_and_wide_head_attention_block():
    """MultiheadAttentionAndNorm on every row with ONE 256-wide head (the width class the flash kernels do not cover) and with
    a 64-wide head (flash path), against the oracle's module."""
    from avssl.module.kw_modules.TransformerModels import MultiheadAttentionAndNorm
    from oracle import speechclip as osc
    for d, nhead in ((256, 1), (64, 1), (128, 2)):
        torch.manual_seed(8)
        m = MultiheadAttentionAndNorm(d_model=d, nhead=nhead)
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn_like(p))
        o = osc.AttentionAndNorm(d_model=d, nhead=nhead).eval()
        o.load_state_dict(m.state_dict())
        B, L = 3, 45
        src = torch.randn(B, L, d)
        lens = torch.tensor([45, 20, 33])
        kpm = torch.arange(L)[None, :] >= lens[:, None]
        ref = o(src.half().float(), kpm)
        m = m.to(DEV)
        got = m(src.to(DEV), kpm.to(DEV))
        valid = ~kpm
        assert (got.cpu()[valid] - ref.detach()[valid]).abs().max() < 1e-2
        hs = m.extract_hidden_states(src.to(DEV), kpm.to(DEV))
        assert len(hs) == 2 and (hs[0].cpu() - src.half().float()).abs().max() == 0


# ---------------------------------------------------------------------------------------------------- text tower
def _clip_pair(name):
    from avssl.module import ClipModel
    from oracle import clip as oc
    m = ClipModel(name)
    o = oc.CLIP(oc.ClipCfg.named(name)).eval()
    missing, unexpected = o.load_state_dict(m.model.state_dict(), strict=False)
    assert not missing, missing
    return m.to(DEV), o


@pytest.mark.parametrize("name,B", [("tiny_c", 6), ("ViT-B/32", 4), ("ViT-L/14", 2)])
def test_text_tower_forward_backward_on_live_positions(name, B):
    """encode_keywords on the K+2 live positions == the oracle's full 77-position evaluation; activation gradients match autograd."""
    from speechclip_b200.functional import workspace
    m, o = _clip_pair(name)
    K, W = 8, m.arch.t_width
    g = torch.Generator().manual_seed(9)
    E = m.model.token_embedding.weight.detach().cpu()
    kw = E[torch.randint(0, E.shape[0], (B, K), generator=g)].clone()
    sot, eot = m.special_tokens()
    kr = kw.clone().requires_grad_(True)
    ref = o.encode_keywords(kr, K, sot, eot)
    dfeat = torch.randn(ref.shape, generator=g)
    ref.backward(dfeat)
    got = m.encode_keywords(kw.to(DEV), K)
    assert rel_err(got.cpu(), ref.detach()) < 1e-2, rel_err(got.cpu(), ref.detach())
    # backward through the plan
    plan = m.text_plan(DEV)
    x0 = torch.empty(B, K + 2, W, device=DEV)
    x0[:, 0] = E[sot].to(DEV)
    x0[:, K + 1] = E[eot].to(DEV)
    x0[:, 1:K + 1] = kw.to(DEV)
    x0 += plan.pos[:K + 2]
    ws = workspace(DEV)
    feat, saved = plan.forward(ws, x0.contiguous(), K + 1, save=True)
    assert rel_err(feat.cpu(), ref.detach()) < 1e-2
    dx0 = plan.backward(ws, saved, dfeat.to(DEV))
    assert rel_err(dx0[:, 1:K + 1].cpu(), kr.grad) < 3e-2, rel_err(dx0[:, 1:K + 1].cpu(), kr.grad)


def test_encode_text_matches_oracle():
    m, o = _clip_pair("tiny_c")
    g = torch.Generator().manual_seed(10)
    B, L = 5, 16
    text = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(3, L - 1, (1,), generator=g))
        text[b, 0] = 94
        text[b, 1:n] = torch.randint(1, 90, (n - 1,), generator=g)
        text[b, n] = 95
    ref = o.encode_text(text)
    got = m.encode_text(text.to(DEV))
    assert rel_err(got.cpu(), ref.detach()) < 1e-2


# ---------------------------------------------------------------------------------------------------- the branch
ZERO_GRAD = ("self_att.attentionBlock_Norm.bias", "linear_proj.bias")


def _check_branch_grads(mine: dict, ref: dict, tol: float, prefix: str = "") -> int:
    """Relative max-error of every trainable gradient.  Train-mode BatchNorm removes any per-feature constant, so the TRUE
    gradients of the two biases feeding it are identically zero: both sides hold rounding noise there, which is bounded
    against the scale of the neighbouring LayerNorm-weight gradient instead of compared relatively."""
    n = 0
    scale = ref[prefix + "self_att.attentionBlock_Norm.weight"].grad.abs().max().item()
    for name, p in mine.items():
        if ".clip." in "." + name or not p.requires_grad or not name.startswith(prefix):
            continue
        og = ref[name].grad
        assert p.grad is not None and og is not None, name
        if name[len(prefix):] in ZERO_GRAD:
            assert og.abs().max().item() < 1e-3 * scale and p.grad.abs().max().item() < 1e-2 * scale, (name, p.grad.abs().max().item(), scale)
        else:
            e = rel_err(p.grad.cpu(), og)
            assert e < tol, (name, e)
        n += 1
    return n


def _branch_pair(size, K=8, seed=0, vocab_npy=None, dropout=0.0):
    """dropout = 0: these tests run the model in train mode for the BatchNorm batch statistics / straight-through path and
    compare with the oracle's deterministic arithmetic; the train-mode attention dropout has its own test (test_dropout_gpu.py)."""
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from speechclip_b200.configs import cascaded_config
    cfg = cascaded_config(size, vocab_npy)
    cfg["model_settings"]["cascaded_branch"]["keyword"]["number"] = K
    cfg["model_settings"]["cascaded_branch"]["transformer_args"]["dropout"] = dropout
    torch.manual_seed(seed)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg))
    with torch.no_grad():
        g = torch.Generator().manual_seed(seed + 1)
        model.audio_encoder.weightedsum_layer.weights.copy_(0.5 * torch.randn(model.audio_encoder.weightedsum_layer.weights.shape, generator=g))
        for n, p in model.cascaded_branch.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "Norm" in n) and "bn_layer" not in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    sot, eot = model.clip.special_tokens()
    ta = cfg["model_settings"]["cascaded_branch"]["transformer_args"]
    ccfg = oc.ClipCfg.named(cfg["clip"]["name"])
    ccfg.vocab = model.clip.model.token_embedding.weight.shape[0]
    oracle = osc.SpeechClipOracle(oh.HubertCfg.named(cfg["audio_encoder"]["name"]), ccfg, None,
                                  dict(temperature=0.07, temperature_trainable=cfg["cl_loss"]["args"]["temperature_trainable"]),
                                  normalize_hiddenstates=cfg["audio_encoder"]["normalize_hiddenstates"],
                                  cascaded_args=dict(keyword_num=K, nhead=ta["nhead"], vq_temp=0.1, sot_token=sot, eot_token=eot)).eval()
    sd = {k: v for k, v in model.state_dict().items() if not k.startswith("cascaded_branch.clip.") and "vector_quantizer" not in k}
    missing, unexpected = oracle.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    oracle.cascaded_branch.bn_layer.train()   # batch statistics, as in training
    return cfg, model.to(DEV), oracle


def test_cascaded_branch_forward_backward_vs_oracle():
    """KW_CascadedBranch.forward on given audio features: BatchNorm output, selected ids (bit-exact), text feature, and the
    gradients of every branch parameter and of the audio features."""
    cfg, model, oracle = _branch_pair("tiny")
    model.train()
    B, T, d, K = 12, 19, 64, 8
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(B, T, d, generator=g)
    lens = torch.randint(8, T + 1, (B,), generator=g)
    lens[0] = T
    fr = feat.half().float().requires_grad_(True)   # the K/V GEMM consumes fp16-rounded features
    collect = {}
    ob = oracle.cascaded_branch
    ofeat, ovq, okw = ob(fr, lens, training=True, collect=collect)
    w = torch.randn(ofeat.shape, generator=g)
    (ofeat * w).sum().backward()

    fd = feat.to(DEV).requires_grad_(True)
    cb = model.cascaded_branch
    nbt0 = int(cb.bn_layer.bn_layer.num_batches_tracked)
    mfeat, mvq, mkw = cb(fd, lens.to(DEV))
    (mfeat * w.to(DEV)).sum().backward()
    assert torch.equal(mvq["targets"].cpu(), ovq["targets"]), "selected vocabulary ids differ"
    E = model.clip.model.token_embedding.weight.detach().cpu()
    assert torch.equal(mkw.cpu(), E[mvq["targets"].cpu().view(B, K)]), "keywords are rows of the table, bit-identical"
    assert (mkw.cpu() - okw.detach()).abs().max() < 1e-6   # oracle: (one-hot + soft - soft.detach()) @ E
    assert rel_err(mfeat.detach().cpu(), ofeat.detach()) < 1e-2
    assert int(cb.bn_layer.bn_layer.num_batches_tracked) == nbt0 + 1
    assert rel_err(cb.bn_layer.bn_layer.running_mean.cpu(), ob.bn_layer.bn_layer.running_mean) < 5e-3
    assert rel_err(cb.bn_layer.bn_layer.running_var.cpu(), ob.bn_layer.bn_layer.running_var) < 5e-3
    assert mvq["temp"] == pytest.approx(0.1) and mvq["num_vars"] == 96
    for k in ("code_perplexity", "prob_perplexity", "diversity_loss"):
        assert abs(mvq[k].item() - ovq[k].item()) < 5e-3 * max(1.0, abs(ovq[k].item())), k
    assert mvq["subword_prob"].shape == (B, K, 96) and torch.equal(mvq["subword_prob"].argmax(-1).cpu(), ovq["targets"].squeeze(-1))
    op = dict(ob.named_parameters())
    n = _check_branch_grads(dict(cb.named_parameters()), op, 4e-2)
    assert n == 11
    assert rel_err(fd.grad.cpu(), fr.grad) < 4e-2
    # padded frames receive no gradient
    for b in range(B):
        assert fd.grad[b, int(lens[b]):].abs().max().item() == 0 if lens[b] < T else True


def test_cascaded_branch_eval_and_hidden_states():
    cfg, model, oracle = _branch_pair("tiny")
    model.eval()
    oracle.cascaded_branch.bn_layer.eval()
    B, T, d = 4, 15, 64
    g = torch.Generator().manual_seed(12)
    feat = torch.randn(B, T, d, generator=g).half().float()
    lens = torch.tensor([15, 9, 12, 15])
    with torch.no_grad():
        ofeat, ovq, okw = oracle.cascaded_branch(feat, lens, training=False)
        mfeat, mvq, mkw = model.cascaded_branch(feat.to(DEV), lens.to(DEV))
    assert torch.equal(mvq["targets"].cpu(), ovq["targets"])
    assert rel_err(mfeat.cpu(), ofeat) < 1e-2
    hs = model.cascaded_branch.extract_hidden_states(feat.to(DEV), lens.to(DEV))
    src = torch.cat([oracle.cascaded_branch.cls.detach().half().float().expand(B, -1, -1), feat], 1)
    kpm = torch.arange(T + 8)[None, :] >= (lens + 8)[:, None]
    ref = oracle.cascaded_branch.self_att(src, kpm)[:, 8:]
    assert len(hs) == 2 and hs[0].shape == (B, T, d)
    valid = ~kpm[:, 8:]
    assert (hs[1].cpu()[valid] - ref.detach()[valid]).abs().max() < 1e-2


# ---------------------------------------------------------------------------------------------------- the model
def _batch(lens, size, seed=3, ids=None):
    g = torch.Generator().manual_seed(seed)
    wavs = [0.1 * torch.randn(n, generator=g) for n in lens]
    B = len(lens)
    img = torch.randn(B, 3, size, size, generator=g)
    ids = torch.arange(B) if ids is None else ids
    padded = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True)
    return wavs, img, ids, {"wav": padded.to(DEV), "wav_len": torch.tensor(lens).to(DEV), "image": img.to(DEV), "id": ids.to(DEV)}


@pytest.mark.parametrize("size", ["tiny", "tiny_large"])
def test_cascaded_model_training_step_vs_oracle(size):
    cfg, model, oracle = _branch_pair(size)
    model.train()
    lens = [6000, 4100, 5555, 6000, 3000, 4800, 6000, 5000, 5200, 4444]
    ids = torch.tensor([0, 1, 1, 2, 3, 3, 4, 5, 6, 7])
    wavs, img, ids, b = _batch(lens, 32, ids=ids)
    out = model.training_step(b)
    feats = out["loss_feats"]
    assert out["log_metrics"]["softmax_temp"] == pytest.approx(0.1)
    loss = model.training_step_end(out)["loss"]
    loss.backward()
    # 1) the oracle on its own: the ids the CUDA path selected are the oracle's, except where the oracle's two best scores
    #    are within the towers' fp16 rounding of each other (the upstream features agree to ~1e-3, not bitwise)
    with torch.no_grad():
        _, _, others = model(b)
        free = oracle(wavs, img, ids)
    mine = others["vq_results"]["targets"].cpu().view(-1)
    ref_scores = free["cascaded_collect"]["cos"].detach().view(mine.numel(), -1).clone()
    ref_scores[:, [0, 2, 3]] = float("-inf")
    _assert_same_ids(mine, free["vq_results"]["targets"].view(-1), ref_scores, tie=5e-3, min_agree=0.9)
    # 2) downstream of the selection, with the oracle held to the same ids: features, loss, every gradient
    oracle.force_idx = mine
    of = oracle(wavs, img, ids)
    oloss = oracle.compute_loss(of)
    oloss.backward()
    assert (feats["cascaded_audio_feat"].detach().cpu() - of["cascaded_audio_feat"].detach()).abs().max() < 1e-2
    assert abs(loss.item() - oloss.item()) < 5e-3 * max(1.0, abs(oloss.item())), (loss.item(), oloss.item())
    oparams = dict(oracle.named_parameters())
    checked = _check_branch_grads({n: p for n, p in model.named_parameters() if n.startswith("cascaded_branch.")}, oparams, 6e-2,
                                  prefix="cascaded_branch.")
    for name, p in model.named_parameters():
        if p.requires_grad and not name.startswith("cascaded_branch."):
            e = rel_err(p.grad.cpu(), oparams[name].grad)
            assert e < 6e-2, (name, e)
            checked += 1
    assert checked == (13 if size == "tiny_large" else 12), checked


def test_cascaded_training_steps_follow_torch_adam():
    """3 fused clip+Adam steps (global-norm clip 4.0, linear warm-up) of the cascaded model track torch.optim.Adam on the
    oracle: loss trajectory and mean parameter update.  The oracle is held to the ids the CUDA path selected at each step."""
    from oracle import speechclip as osc
    cfg, model, oracle = _branch_pair("tiny")
    model.train()
    cfg["audio_encoder"]["scheduler"]["warmup"] = 2
    model.config.audio_encoder.scheduler.warmup = 2
    opts, scheds = model.configure_optimizers()
    opt, sched = opts[0], scheds[0]["scheduler"]
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    start = {n: p.detach().cpu().clone() for n, p in model.named_parameters() if p.requires_grad}
    oparams = [p for n, p in oracle.named_parameters() if n in trainable]
    oopt = torch.optim.Adam(oparams, lr=1e-4, weight_decay=1e-6)
    losses, olosses = [], []
    for step in range(3):
        wavs, img, ids, b = _batch([5000, 6000, 4000, 6000, 5500, 4700], 32, seed=20 + step)
        losses_, log_metrics, others = model(b)
        loss = model.training_step_end({"loss_feats": losses_, "log_metrics": log_metrics})["loss"]
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        losses.append(loss.item())
        oracle.force_idx = others["vq_results"]["targets"].cpu().view(-1)
        ol = oracle.compute_loss(oracle(wavs, img, ids))
        oopt.zero_grad()
        ol.backward()
        torch.nn.utils.clip_grad_norm_(oparams, 4.0)
        for gparam in oopt.param_groups:
            gparam["lr"] = 1e-4 * osc.linear_warmup_decay(step, 1e-4, 2, 50000, 1e-8)
        oopt.step()
        olosses.append(ol.item())
    for a, r in zip(losses, olosses):
        assert abs(a - r) < 5e-3 * max(1.0, abs(r)), (losses, olosses)
    od = dict(oracle.named_parameters())
    for n, p in model.named_parameters():
        if not p.requires_grad or n[len("cascaded_branch."):] in ZERO_GRAD:
            continue   # zero true gradient: both sides step by sign(noise) * lr
        mine, ref = p.detach().cpu() - start[n], od[n].detach() - start[n]
        if n.endswith("in_proj_bias"):
            d3 = mine.numel() // 3   # the key bias has zero true gradient as well (shifts every score of a row equally)
            keep = torch.cat([torch.arange(0, d3), torch.arange(2 * d3, 3 * d3)])
            mine, ref = mine[keep], ref[keep]
        assert ref.abs().mean() > 1e-5, n
        assert (mine - ref).abs().mean() < 0.15 * ref.abs().mean(), (n, (mine - ref).abs().mean().item(), ref.abs().mean().item())
    assert model.arena().intact()


def test_cascaded_full_size_base_step_runs():
    """Cascaded-base shapes (HuBERT-base, ViT-B/32, 8112-entry reduced vocabulary, K = 8) run one training step; the selected
    ids and keywords agree with an fp32 re-evaluation of the quantiser on the branch's own BatchNorm output."""
    import os
    import tempfile
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200.configs import cascaded_config, write_synthetic_vocab_usage
    with tempfile.TemporaryDirectory() as td:
        npy = write_synthetic_vocab_usage(os.path.join(td, "usage.npy"))
        torch.manual_seed(0)
        model = KWClip_GeneralTransformer(OrderedNamespace(cascaded_config("base", npy))).to(DEV).train()
    B = 16
    g = torch.Generator().manual_seed(13)
    b = {"wav": (0.1 * torch.randn(B, 40000, generator=g)).to(DEV), "wav_len": torch.full((B,), 40000).to(DEV),
         "image": torch.randn(B, 3, 224, 224, generator=g).to(DEV), "id": torch.arange(B).to(DEV)}
    opts, _ = model.configure_optimizers()
    out = model.training_step(b)
    loss = model.training_step_end(out)["loss"]
    loss.backward()
    assert math.isfinite(loss.item())
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    opts[0].step()
    _, _, others = model(b)
    assert others["keywords"].shape == (B, 8, 512) and others["vq_results"]["targets"].shape == (B, 8, 1)
    assert not torch.isin(others["vq_results"]["targets"], torch.tensor([0, 2, 3], device=DEV)).any()
