"""Generate the golden fixtures that pin ``oracle/``.  Run in the BUILD container only:

    python tests/golden/make_golden.py

Part 1 imports the reference's own torch-only source files straight from /root/reference
(read-only; nothing is copied) and records their outputs on seeded inputs:
``ref_*.npz``.  Part 2 records outputs of the independent in-container implementations of the
third-party towers (``transformers`` Hubert / CLIP) on weights produced by
``speechclip_b200.init.seeded_init_``: ``hf_*.npz``.  /root/reference does not exist on the GPU
box, so tests only ever read the committed ``.npz`` files.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/avssl"


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: out[k].shape for k in out})


def part1_reference():
    g = torch.Generator().manual_seed(7122)
    # ---- MaskedContrastiveLoss (losses.py:129-245) incl. autograd grads
    losses = load_ref("module/losses.py", "ref_losses")
    for tag, B, D in (("small", 16, 32), ("mid", 64, 128)):
        a = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1).requires_grad_()
        b = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1).requires_grad_()
        ids = torch.randperm(B, generator=g) // 3            # same-id groups exercise the negative mask
        crit = losses.MaskedContrastiveLoss(temperature=0.07)
        loss = crit(feat_A=a, feat_B=b, index=ids)
        loss.backward()
        crit_t = losses.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True)
        a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
        loss_t = crit_t(feat_A=a2, feat_B=b2, index=ids)
        loss_t.backward()
        loss_noid = crit(feat_A=a.detach(), feat_B=b.detach(), index=None)
        save(f"ref_loss_{tag}.npz", a=a, b=b, ids=ids, loss=loss, da=a.grad, db=b.grad,
             loss_t=loss_t, da_t=a2.grad, db_t=b2.grad, dtemp=crit_t.temperature.grad,
             temp_param=crit_t.temperature, loss_noid=loss_noid)
    try:
        losses.MaskedContrastiveLoss()(torch.randn(300, 8), torch.randn(300, 8), torch.arange(300))
        raised = False
    except IndexError:
        raised = True
    print("reference loss raises IndexError for B>256 (MAX_EYE):", raised)

    # ---- WeightedSumLayer (weighted_sum.py:26-45)
    ws = load_ref("module/weighted_sum.py", "ref_ws")
    hidden = [torch.randn(2, 7, 16, generator=g) for _ in range(5)]
    w = torch.randn(5, generator=g)
    outs = {}
    for norm in (False, True):
        layer = ws.WeightedSumLayer(5, normalize_features=norm)
        layer.weights.data.copy_(w)
        outs[f"out_norm{int(norm)}"] = layer(hidden)
    save("ref_weighted_sum.npz", hidden=torch.stack(hidden), weights=w, **outs)

    # ---- get_keypadding_mask (data_utils.py:4-20)
    du = load_ref("util/data_utils.py", "ref_du")
    lens = torch.tensor([1, 5, 9, 12, 3])
    save("ref_keypad.npz", lens=lens, mask=du.get_keypadding_mask(12, lens))

    # ---- TransformerEncoder branch (TransformerModels.py:48-96), eval mode
    tm = load_ref("module/kw_modules/TransformerModels.py", "ref_tm")
    for tag, norm_first in (("postln", False), ("preln", True)):
        torch.manual_seed(11)
        enc = tm.TransformerEncoder(n_layers=1, d_model=64, nhead=8, dim_feedforward=128, dropout=0.1,
                                    activation="gelu", layer_norm_eps=1e-5, batch_first=True, norm_first=norm_first)
        for p in enc.parameters():
            p.data.add_(0.05 * torch.randn(p.shape, generator=g))
        enc.eval()
        src = torch.randn(3, 11, 64, generator=g)
        kpm = du.get_keypadding_mask(11, torch.tensor([11, 4, 8]))
        with torch.no_grad():
            out = enc(src, kpm)
            hs = enc.extract_hidden_states(src, kpm)
        sd = {"sd." + k: v for k, v in enc.state_dict().items()}
        save(f"ref_branch_{tag}.npz", src=src, kpm=kpm, out=out, h0=hs[0], h1=hs[1], **sd)

    # ---- mutualRetrieval (retrieval.py:6-121)
    rt = load_ref("module/retrieval.py", "ref_rt")
    nA, nB = 40, 8
    score = torch.randn(nA, nB, generator=g)
    ab = torch.arange(nA) // 5
    ba = torch.arange(nB)
    rAB, rBA, rM = rt.mutualRetrieval(score, score.T.contiguous(), ab, ba, [1, 5, 10])
    save("ref_retrieval.npz", score=score, ab=ab, ba=ba,
         rAB=[rAB[f"recall@{k}"] for k in (1, 5, 10)], rBA=[rBA[f"recall@{k}"] for k in (1, 5, 10)],
         rM=[rM[f"recall@{k}"] for k in (1, 5, 10)])

    # ---- linear_warmup_decay (scheduler.py:22-38)
    sch = load_ref("optim/scheduler.py", "ref_sch")
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    s = sch.get_scheduler("linear_warmup_decay", opt, warmup=5, max_step=20, final_lr=1e-8)
    lrs = []
    for _ in range(20):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        s.step()
    save("ref_scheduler.npz", lrs=np.array(lrs, dtype=np.float64))

    # ---- cascaded-branch pieces: MultiheadAttentionAndNorm (TransformerModels.py:99-135), Kw_BatchNorm (kw_bn.py:8-164,
    #      eachKw / parallel), SimpleVectorQuantizer (my_vector_quantizer.py:12-165, fixed temperature, hard straight-through)
    torch.manual_seed(13)
    mha = tm.MultiheadAttentionAndNorm(d_model=64, nhead=1, dropout=0.1)
    for p in mha.parameters():
        p.data.add_(0.05 * torch.randn(p.shape, generator=g))
    mha.eval()
    src = torch.randn(3, 12, 64, generator=g)
    kpm = du.get_keypadding_mask(12, torch.tensor([12, 6, 9]))
    with torch.no_grad():
        out = mha(src, kpm)
    save("ref_mha_norm.npz", src=src, kpm=kpm, out=out, **{"sd." + k: v for k, v in mha.state_dict().items()})

    kb = load_ref("module/speechclip_c_modules/kw_bn.py", "ref_kb")
    ib, isc = torch.randn(16, generator=g), torch.rand(16, generator=g) + 0.5
    bn = kb.Kw_BatchNorm(kw_num=4, kw_dim=16, batchnorm_type="eachKw", init_bias=ib, init_scale=isc, std_scale=1.0,
                         learnable=True, parallel=True).train()
    x = torch.randn(6, 4, 16, generator=g).requires_grad_()
    y = bn(x)
    w = torch.randn(6, 4, 16, generator=g)
    (y * w).sum().backward()
    bn.eval()
    with torch.no_grad():
        y_eval = bn(x.detach())
    save("ref_kw_bn.npz", x=x, init_bias=ib, init_scale=isc, y_train=y, w=w, dx=x.grad, dgamma=bn.bn_layer.weight.grad,
         dbeta=bn.bn_layer.bias.grad, running_mean=bn.bn_layer.running_mean, running_var=bn.bn_layer.running_var, y_eval=y_eval)

    vqm = load_ref("module/speechclip_c_modules/my_vector_quantizer.py", "ref_vq")
    vq = vqm.SimpleVectorQuantizer(temp="fixed=0.1", time_first=True, use_gumbel=False, hard=True).train()
    cos = (0.3 * torch.randn(3, 4, 40, generator=g)).requires_grad_()
    r = vq(cos.clone())
    wv = torch.randn(3, 4, 40, generator=g)
    (r["subword_prob"] * wv).sum().backward()
    vq.eval()
    with torch.no_grad():
        r_eval = vq(cos.detach().clone())
    save("ref_vq.npz", cos=cos, w=wv, dcos=cos.grad, subword_prob=r["subword_prob"], targets=r["targets"],
         code_perplexity=r["code_perplexity"], prob_perplexity=r["prob_perplexity"], ent_per_t=r["ent_per_t"],
         diversity_loss=r["diversity_loss"], subword_prob_eval=r_eval["subword_prob"])

    # ---- TRAIN mode of the two branch encoders (dropout 0.1: spchclp_p.yaml:27, TransformerModels.py:55-75,110-117).  torch draws
    #      its dropout masks from its own generator inside library code, which no other implementation can replay; so the masks are
    #      drawn HERE (from g) and injected into the reference's own modules by temporarily replacing the two library entry points
    #      that consume randomness (F.dropout; F.scaled_dot_product_attention, evaluated as softmax(qk^T/sqrt(d) + mask) * m @ v —
    #      checked bit-exact against torch's CPU kernel, which applies the mask the same way).  Fixtures hold inputs, masks, outputs.
    import torch.nn.functional as F
    drawn = []

    def draw(shape, p):
        m = (torch.rand(shape, generator=g) >= p).float() / (1.0 - p)
        drawn.append(m)
        return m

    def patched_dropout(x, p=0.5, training=True, inplace=False):
        return x * draw(x.shape, p) if training and p > 0 else x

    def patched_sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        sc = q @ k.transpose(-1, -2) * (q.shape[-1] ** -0.5 if scale is None else scale)
        if attn_mask is not None:
            sc = sc.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else sc + attn_mask
        pr = torch.softmax(sc, -1)
        if dropout_p > 0:
            pr = pr * draw(pr.shape, dropout_p)
        return pr @ v

    orig = (F.dropout, F.scaled_dot_product_attention)
    F.dropout, F.scaled_dot_product_attention = patched_dropout, patched_sdpa
    try:
        torch.manual_seed(17)
        enc = tm.TransformerEncoder(n_layers=1, d_model=64, nhead=8, dim_feedforward=128, dropout=0.1, activation="gelu",
                                    layer_norm_eps=1e-5, batch_first=True, norm_first=False)
        for p in enc.parameters():
            p.data.add_(0.05 * torch.randn(p.shape, generator=g))
        enc.train()
        src = torch.randn(4, 11, 64, generator=g)
        kpm = du.get_keypadding_mask(11, torch.tensor([11, 4, 8, 10]))
        drawn.clear()
        out = enc(src, kpm)
        assert [tuple(m.shape) for m in drawn] == [(4, 8, 11, 11), (4, 11, 64), (4, 11, 128), (4, 11, 64)], [m.shape for m in drawn]
        save("ref_branch_train_dropout.npz", src=src, kpm=kpm, out=out, m_attn=drawn[0], m_dropout1=drawn[1], m_ffn=drawn[2],
             m_dropout2=drawn[3], **{"sd." + k: v for k, v in enc.state_dict().items()})

        torch.manual_seed(19)
        mha = tm.MultiheadAttentionAndNorm(d_model=64, nhead=1, dropout=0.1)
        for p in mha.parameters():
            p.data.add_(0.05 * torch.randn(p.shape, generator=g))
        mha.train()
        src = torch.randn(3, 12, 64, generator=g)
        kpm = du.get_keypadding_mask(12, torch.tensor([12, 6, 9]))
        drawn.clear()
        out = mha(src, kpm)
        assert [tuple(m.shape) for m in drawn] == [(3, 12, 12)], [m.shape for m in drawn]   # [B * heads, L, L]
        save("ref_mha_norm_train_dropout.npz", src=src, kpm=kpm, out=out, m_attn=drawn[0].view(3, 1, 12, 12),
             **{"sd." + k: v for k, v in mha.state_dict().items()})
    finally:
        F.dropout, F.scaled_dot_product_attention = orig

    # ---- pooling layers (pooling.py:8-390) and MLPLayers (projections.py:6-29): the reference's own modules on seeded inputs
    pl = load_ref("module/pooling.py", "ref_pool")
    torch.manual_seed(23)
    mp = pl.MeanPoolingLayer(in_dim=24, out_dim=16)
    x = torch.randn(5, 9, 24, generator=g).requires_grad_()
    xl = torch.tensor([9, 3, 7, 1, 5])
    y = mp(x, xl)
    wy = torch.randn(5, 16, generator=g)
    (y * wy).sum().backward()
    y_nolen = mp(x.detach())
    ap = pl.AttentivePoolingLayer(dim_A=12, dim_B=10)
    with torch.no_grad():
        ap.U.copy_(0.3 * torch.randn(12, 10, generator=g))
        a_in, b_in = torch.randn(4, 12, 7, generator=g), torch.randn(4, 10, 5, generator=g)
        msk = ap.generate_input_msk(input_A_lens=torch.tensor([7, 4, 6, 2]), input_B_lens=torch.tensor([5, 5, 3, 1]), max_Alen=7, max_Blen=5)
        oa, ob = ap(a_in, b_in, msk)
        oa_nm, ob_nm = ap(a_in, b_in)
        msk_a = ap.generate_input_msk(input_A_lens=torch.tensor([7, 4, 6, 2]), max_Alen=7, max_Blen=1)
        b2 = torch.randn(3, 10, 5, generator=g)
        boa, bob = ap.batch_forward(a_in, b2, msk_a)
        bvec = torch.randn(10, 6, generator=g)
        emb = ap.cal_batch_embedding(a_in, bvec, msk_a)
    save("ref_pooling.npz", mp_x=x, mp_len=xl, mp_y=y, mp_w=wy, mp_dx=x.grad, mp_y_nolen=y_nolen,
         **{"mp_sd." + k: v for k, v in mp.state_dict().items()}, **{"mp_grad." + k: v.grad for k, v in mp.named_parameters()},
         ap_U=ap.U, ap_a=a_in, ap_b=b_in, ap_msk=msk, ap_oa=oa, ap_ob=ob, ap_oa_nomask=oa_nm, ap_ob_nomask=ob_nm, ap_msk_a=msk_a,
         ap_b2=b2, ap_boa=boa, ap_bob=bob, ap_bvec=bvec, ap_emb=emb)
    pj = load_ref("module/projections.py", "ref_proj")
    torch.manual_seed(29)
    mlp = pj.MLPLayers(units=[16, 32, 8], dropout=0.1).eval()
    xm = torch.randn(6, 16, generator=g).requires_grad_()
    ym = mlp(xm)
    wm = torch.randn(6, 8, generator=g)
    (ym * wm).sum().backward()
    save("ref_mlp.npz", x=xm, y=ym, w=wm, dx=xm.grad, **{"sd." + k: v for k, v in mlp.state_dict().items()},
         **{"grad." + k: v.grad for k, v in mlp.named_parameters()})

    # ---- collate_general (collate_function.py:7-36)
    cf = load_ref("data/collate_function.py", "ref_cf")
    rows = [{"wav": torch.randn(n, generator=g), "image": torch.randn(3, 4, 4, generator=g), "id": i * 3, "text": torch.arange(5).view(1, 5) + i}
            for i, n in enumerate((11, 7, 15))]
    col = cf.collate_general(rows)
    save("ref_collate.npz", **{f"row{i}_{k}": v for i, r in enumerate(rows) for k, v in r.items() if isinstance(v, torch.Tensor)},
         **{"out_" + k: v for k, v in col.items()}, out_keys=np.array(list(col.keys())))

    # ---- random_crop_max_length semantics (audio_transforms.py:5-23): shapes only (np.random offset)
    at = load_ref("data/audio_transforms.py", "ref_at")
    assert at.random_crop_max_length(torch.arange(10), 4, 10).shape == (4,)
    assert at.random_crop_max_length(torch.arange(10), 20, 10).shape == (10,)
    assert at.random_crop_max_length(torch.arange(10), -1, 10).shape == (10,)


def part2_hf():
    import transformers
    from oracle import clip as oc
    from oracle import hubert as oh
    from speechclip_b200.init import seeded_init_
    from tests.hf_map import hf_clip_from_oracle, hf_hubert_from_oracle

    g = torch.Generator().manual_seed(7122)
    for name in ("tiny", "tiny_large"):
        cfg = oh.HubertCfg.named(name)
        om = seeded_init_(oh.HubertModel(cfg), 7122).eval()
        hf = hf_hubert_from_oracle(om)
        wav = 0.1 * torch.randn(2, 4000, generator=g)
        with torch.no_grad():
            o = hf(wav, output_hidden_states=True)
        save(f"hf_hubert_{name}.npz", wav=wav, **{f"h{i}": h for i, h in enumerate(o.hidden_states)})
    ccfg = oc.ClipCfg.named("tiny")
    om = seeded_init_(oc.CLIP(ccfg), 7122).eval()
    hv, ht = hf_clip_from_oracle(om)
    img = torch.randn(3, 3, ccfg.image_size, ccfg.image_size, generator=g)
    tok = torch.randint(1, ccfg.vocab - 1, (3, ccfg.context), generator=g)
    tok[:, 9] = ccfg.vocab - 1  # EOT = largest id, as in CLIP's tokenizer
    with torch.no_grad():
        save("hf_clip_tiny.npz", img=img, tok=tok, image_embeds=hv(pixel_values=img).image_embeds,
             text_embeds=ht(input_ids=tok).text_embeds)


if __name__ == "__main__":
    part1_reference()
    part2_hf()
