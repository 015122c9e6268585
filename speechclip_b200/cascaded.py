"""The cascaded branch of SpeechCLIP, forward AND backward, as kernel sequences over the C ABI.

Reference: ``avssl/model/kwClip.py:697-916`` (KW_CascadedBranch), ``avssl/module/kw_modules/TransformerModels.py:99-135``
(MultiheadAttentionAndNorm), ``avssl/module/speechclip_c_modules/kw_bn.py:96-125`` (Kw_BatchNorm, eachKw/parallel),
``avssl/module/speechclip_c_modules/my_vector_quantizer.py:66-135`` (SimpleVectorQuantizer) and
``avssl/module/clip_official.py:220-268`` (ClipModel.encode_keywords -> frozen CLIP text transformer).

What the reference computes, and what is pruned here without changing the result:

* ``self_att(src)[:, :K]``: only the K keyword rows of the attention block are consumed.  Their queries are the learned
  [CLS] vectors (identical for every utterance), so the block is one K/V GEMM over all rows, a K-query attention per
  utterance (scb_mq_attention_fwd) and [B*K, d] row operations; the out-proj / LayerNorm of the B*T frame rows, which the
  reference computes and discards, are never evaluated.
* ``clip.encode_keywords``: the text transformer is causal and only position K+1 (the [EOT] slot) is read, so positions
  beyond K+1 (67 of 77) cannot influence the result and are not computed.
* ``subword_prob @ token_embedding``: the forward VALUE of the hard straight-through estimator is a one-hot row, so the
  product is a gather; the backward is the dense softmax(cos / temp) Jacobian exactly as autograd would apply it.

The keyword-vs-vocabulary scores whose argmax selects the tokens are computed on the tensor cores with the 3xTF32 split
(scb_split_tf32) so that the selected indices match an fp32 evaluation.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .engine import H, Workspace, _f32, _h
from .head import _new, dgrad, linear, transposed, wgrad, _ceil4

BF = torch.bfloat16
A0 = "self_att.multihead_attn_layer."
PARAM_ORDER = ["cls", A0 + "in_proj_weight", A0 + "in_proj_bias", A0 + "out_proj.weight", A0 + "out_proj.bias",
               "self_att.attentionBlock_Norm.weight", "self_att.attentionBlock_Norm.bias", "linear_proj.weight", "linear_proj.bias",
               "bn_layer.bn_layer.weight", "bn_layer.bn_layer.bias"]
VQ_MASK_IDS = (0, 2, 3)   # prob_msk default of SimpleVectorQuantizer.forward (my_vector_quantizer.py:66)


# ====================================================================================================== CLIP text tower
class TextTowerPlan:
    """Frozen CLIP text transformer (pre-LN, QuickGELU, causal) on the first L positions, with an activation-gradient backward.

    Forward GEMMs run in fp16 like the other frozen towers; the backward carries fp32 gradients through TF32 GEMMs against
    fp32 transposed copies of the (frozen) weights, so no per-step weight transposes or gradient casts are needed."""

    def __init__(self, sd: Dict[str, torch.Tensor], dev, *, heads: int, need_backward: bool = True):
        self.dev, self.heads = dev, heads
        self.pos = _f32(sd["positional_embedding"], dev)
        self.context, self.d = self.pos.shape
        self.layers: List[dict] = []
        l = 0
        while f"transformer.resblocks.{l}.attn.in_proj_weight" in sd:
            p = f"transformer.resblocks.{l}."
            L = dict(wqkv=_h(sd[p + "attn.in_proj_weight"], dev), bqkv=_f32(sd[p + "attn.in_proj_bias"], dev),
                     wo=_h(sd[p + "attn.out_proj.weight"], dev), bo=_f32(sd[p + "attn.out_proj.bias"], dev),
                     ln1=(_f32(sd[p + "ln_1.weight"], dev), _f32(sd[p + "ln_1.bias"], dev)),
                     w1=_h(sd[p + "mlp.c_fc.weight"], dev), b1=_f32(sd[p + "mlp.c_fc.bias"], dev),
                     w2=_h(sd[p + "mlp.c_proj.weight"], dev), b2=_f32(sd[p + "mlp.c_proj.bias"], dev),
                     ln2=(_f32(sd[p + "ln_2.weight"], dev), _f32(sd[p + "ln_2.bias"], dev)))
            if need_backward:
                for name, key in (("wqkv_t", "attn.in_proj_weight"), ("wo_t", "attn.out_proj.weight"), ("w1_t", "mlp.c_fc.weight"),
                                  ("w2_t", "mlp.c_proj.weight")):
                    w = _f32(sd[p + key], dev)
                    L[name] = ops.transpose(w, torch.empty(w.shape[1], w.shape[0], device=dev, dtype=torch.float32))
            self.layers.append(L)
            l += 1
        self.ln_final = (_f32(sd["ln_final.weight"], dev), _f32(sd["ln_final.bias"], dev))
        proj = _f32(sd["text_projection"], dev)                    # [width, embed]
        self.proj = proj                                            # right operand of the backward GEMM as stored
        self.proj_t = ops.transpose(proj, torch.empty(proj.shape[1], proj.shape[0], device=dev, dtype=torch.float32))
        self.embed = proj.shape[1]
        self.need_backward = need_backward

    def forward(self, ws: Workspace, x0: torch.Tensor, last_row, save: bool):
        """x0 fp32 [B, L, d] (token + positional embeddings).  ``last_row``: int (same row for every sequence) or an int64
        CUDA tensor [B] of per-sequence rows.  -> (features fp32 [B, embed], saved activations or None)."""
        B, L, d = x0.shape
        M, dev, heads = B * L, x0.device, self.heads
        hd = d // heads
        x = x0.view(M, d)
        saved = [] if save else None
        a16 = ws.view("tt_a16", (M, d), H)
        ctx = ws.view("tt_ctx", (B, L, d), H)
        act = ws.view("tt_act", (M, 4 * d), H)
        for li, lw in enumerate(self.layers):
            if save:
                st1, st2 = _new((M, 2), dev), _new((M, 2), dev)
                qkv, xa, pre, nxt = _new((B, L, 3 * d), dev, H), _new((M, d), dev), _new((M, 4 * d), dev, H), _new((M, d), dev)
            else:
                st1 = st2 = None
                qkv, xa, pre = ws.view("tt_qkv", (B, L, 3 * d), H), ws.view("tt_xa", (M, d), torch.float32), ws.view("tt_pre", (M, 4 * d), H)
                nxt = ws.view(f"tt_x{li & 1}", (M, d), torch.float32)
            ops.layernorm(x, *lw["ln1"], y16=a16, stats=st1, rows=M, d=d)
            ops.gemm(a16, lw["wqkv"], bias=lw["bqkv"], out=qkv.view(M, 3 * d))
            ops.attention(qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:3 * d], ctx, heads, hd ** -0.5, None, True)
            ops.gemm(ctx.view(M, d), lw["wo"], bias=lw["bo"], residual=x, out=xa)
            ops.layernorm(xa, *lw["ln2"], y16=a16, stats=st2, rows=M, d=d)
            ops.gemm(a16, lw["w1"], bias=lw["b1"], out=pre)
            ops.act16_fwd(pre, ops.ACT_QUICK_GELU, act)
            ops.gemm(act, lw["w2"], bias=lw["b2"], residual=xa, out=nxt)
            if save:
                saved.append(dict(x=x, st1=st1, qkv=qkv, xa=xa, st2=st2, pre=pre))
            x = nxt
        xl = _new((B, d), dev)
        if isinstance(last_row, int):
            ops.cast_rows(x.view(B, L * d)[:, last_row * d:(last_row + 1) * d], xl, rows=B, cols=d, src_ld=L * d, dst_ld=d)
        else:
            ops.gather_rows(x.view(B, L, d), last_row, xl)
        y, stf = _new((B, d), dev), _new((B, 2), dev)
        ops.layernorm(xl, *self.ln_final, y32=y, stats=stf, rows=B, d=d)
        feat = _new((B, self.embed), dev)
        linear(y, self.proj_t, feat)
        if save:
            return feat, dict(layers=saved, xl=xl, stf=stf, B=B, L=L, last_row=last_row)
        return feat, None

    def backward(self, ws: Workspace, s: dict, dfeat: torch.Tensor) -> torch.Tensor:
        """dfeat fp32 [B, embed] -> d x0 fp32 [B, L, d]."""
        assert self.need_backward and isinstance(s["last_row"], int)
        B, L, d, dev, heads = s["B"], s["L"], self.d, dfeat.device, self.heads
        M, hd = B * L, self.d // self.heads
        dy = _new((B, d), dev)
        linear(dfeat, self.proj, dy)                                  # dy = dfeat @ proj^T  (proj is [d, embed] = [N, K])
        dxl = _new((B, d), dev)
        ops.layernorm_bwd(dy, s["xl"], s["stf"], self.ln_final[0], dxl, None, None)
        dx = _new((B, L, d), dev, zero=True)
        r = s["last_row"]
        ops.cast_rows(dxl, dx.view(B, L * d)[:, r * d:(r + 1) * d], rows=B, cols=d, src_ld=d, dst_ld=L * d)
        dx = dx.view(M, d)
        dact = ws.view("tt_dact", (M, 4 * d), torch.float32)
        dln = ws.view("tt_dln", (M, d), torch.float32)
        dctx = ws.view("tt_dctx", (M, d), torch.float32)
        dqkv = ws.view("tt_dqkv", (M, 3 * d), torch.float32)
        dtmp = ws.view("tt_dtmp", (M, d), torch.float32)
        for lw, a in zip(reversed(self.layers), reversed(s["layers"])):
            # MLP: x_out = xa + W2 act(W1 LN2(xa) + b1) + b2
            linear(dx, lw["w2_t"], dact)                              # [M, 4d] = dx @ W2   (w2_t = W2^T is [4d, d] = [N, K])
            ops.act_bwd(dact, a["pre"], ops.ACT_QUICK_GELU, dact)
            linear(dact, lw["w1_t"], dln)                             # [M, d] = dpre @ W1
            ops.layernorm_bwd(dln, a["xa"], a["st2"], lw["ln2"][0], dtmp, None, None)
            ops.rows_bias_act(dtmp, None, dx, d, ops.ACT_NONE, None, dx)   # dxa = dx (residual) + LN2 path
            # attention: xa = x + Wo attn(LN1(x)) + bo
            linear(dx, lw["wo_t"], dctx)
            ops.attention_small_bwd(a["qkv"], dctx, dqkv, B, L, heads, hd, hd ** -0.5, True)
            linear(dqkv, lw["wqkv_t"], dln)
            ops.layernorm_bwd(dln, a["x"], a["st1"], lw["ln1"][0], dtmp, None, None)
            ops.rows_bias_act(dtmp, None, dx, d, ops.ACT_NONE, None, dx)
        return dx.view(B, L, d)


# ====================================================================================================== vocabulary
class Vocabulary:
    """Frozen token-embedding table prepared for the cosine / VQ kernels (built once per device)."""

    def __init__(self, emb: torch.Tensor, dev):
        E = _f32(emb, dev)
        self.E = E                                                      # [V, W]
        self.V, self.W = E.shape
        self.norm = torch.linalg.vector_norm(E, dim=1).contiguous()    # one-off constant preparation
        self.split = ops.split_tf32(E, torch.empty(self.V, 3 * self.W, device=dev, dtype=torch.float32), 1)
        Vp = _ceil4(self.V)
        # unit rows, transposed: right operand [N = W][K = V] of the cosine backward GEMM
        self.unit_t = torch.zeros(self.W, Vp, device=dev, dtype=torch.float32)
        ops.transpose((E / self.norm.clamp_min(1e-20)[:, None]).contiguous(), self.unit_t[:, :self.V])
        self.mask_ids = torch.tensor(VQ_MASK_IDS, device=dev, dtype=torch.int32)


# ====================================================================================================== the branch
SITE_MQ_ATTN = 8   # dropout site of the keyword attention weights: element ((b * heads + h) * K + k) * (T + K) + j


class CascadedHead:
    """Stateless executor of KW_CascadedBranch.forward; ``p`` maps PARAM_ORDER names to live fp32 CUDA tensors."""

    def __init__(self, d_model: int, nhead: int, keyword_num: int, text_dim: int, eps: float = 1e-5, bn_eps: float = 1e-5,
                 bn_momentum: float = 0.1):
        self.d, self.heads, self.K, self.W = d_model, nhead, keyword_num, text_dim
        self.hd = d_model // nhead
        self.eps, self.bn_eps, self.bn_momentum = eps, bn_eps, bn_momentum

    def _build_src(self, p, audio_feat: torch.Tensor) -> torch.Tensor:
        B, T, d = audio_feat.shape
        K, Tk = self.K, T + self.K
        src = _new((B, Tk, d), audio_feat.device, H)
        ops.cast_rows(audio_feat, src.view(B, Tk * d)[:, K * d:], rows=B, cols=T * d, src_ld=T * d, dst_ld=Tk * d)
        ops.broadcast_row(p["cls"].view(-1), None, src, Tk * d, B, K * d)
        return src

    def keywords_forward(self, ws: Workspace, p, audio_feat: torch.Tensor, kv_len: torch.Tensor, bn_buffers, training: bool, drop=None):
        """audio_feat fp32 [B, T, d]; kv_len int32 [B] = audio_len + K.  -> (kw_bn fp32 [B, K, W], saved).
        kwClip.py:866-887: attention block on the keyword rows, linear_proj, Kw_BatchNorm.
        drop = (p, rng_state): train-mode attention dropout of nn.MultiheadAttention (TransformerModels.py:110-117)."""
        B, T, d = audio_feat.shape
        drop = (float(drop[0]), drop[1].clone(), SITE_MQ_ATTN) if drop is not None and drop[0] > 0 else None
        K, W, hd, heads, dev = self.K, self.W, self.hd, self.heads, audio_feat.device
        Tk, R = T + K, B * K
        M = B * Tk
        w_in, b_in = p[A0 + "in_proj_weight"], p[A0 + "in_proj_bias"]
        src = self._build_src(p, audio_feat)
        wkv16 = ws.view("casc_wkv16", (2 * d, d), H)
        ops.cast_rows(w_in[d:], wkv16)
        kv = _new((B, Tk, 2 * d), dev, H)
        ops.gemm(src.view(M, d), wkv16, bias=b_in[d:], out=kv.view(M, 2 * d))
        cls = p["cls"].view(K, d)
        q = _new((K, d), dev)
        ops.sgemm(cls, w_in[:d], q)
        ops.rows_bias_act(q, b_in[:d], None, 0, ops.ACT_NONE, None, q)
        probs = _new((B, heads, K, Tk), dev)
        ctx = _new((B, K, d), dev)
        ops.mq_attention_fwd(q, kv, 0, d, kv_len, heads, hd, hd ** -0.5, probs, ctx, drop=drop)
        cls_rows = ws.view("casc_cls_rows", (R, d), torch.float32)
        ops.broadcast_row(p["cls"].view(-1), None, cls_rows, K * d, B, K * d)
        t1 = _new((R, d), dev)
        linear(ctx.view(R, d), p[A0 + "out_proj.weight"], t1, bias=p[A0 + "out_proj.bias"], residual=cls_rows)
        x1, st1 = _new((R, d), dev), _new((R, 2), dev)
        ops.layernorm(t1, p["self_att.attentionBlock_Norm.weight"], p["self_att.attentionBlock_Norm.bias"], y32=x1, stats=st1, eps=self.eps)
        kwp = _new((B, K, W), dev)
        linear(x1, p["linear_proj.weight"], kwp.view(R, W), bias=p["linear_proj.bias"])
        kw_bn = _new((B, K, W), dev)
        mean, rstd = (_new((K * W,), dev), _new((K * W,), dev)) if training else (None, None)
        rm, rv = bn_buffers
        ops.batchnorm_fwd(kwp, kw_bn, p["bn_layer.bn_layer.weight"], p["bn_layer.bn_layer.bias"], rm, rv, mean, rstd, self.bn_eps,
                          self.bn_momentum, training)
        saved = dict(B=B, T=T, src=src, kv=kv, q=q, probs=probs, ctx=ctx, t1=t1, x1=x1, st1=st1, kwp=kwp, mean=mean, rstd=rstd,
                     kv_len=kv_len, kw_bn=kw_bn, drop=drop)
        return kw_bn, saved

    def quantize(self, ws: Workspace, vocab: Vocabulary, kw_bn: torch.Tensor, temp: float):
        """Cosine scores against the vocabulary + SimpleVectorQuantizer.  -> (cos fp32 [R, V] with masked ids = -inf,
        idx int64 [B, K], stats fp32 [R, 4])."""
        B, K, W = kw_bn.shape
        R, dev = B * K, kw_bn.device
        a3 = ws.view("casc_kw_split", (R, 3 * W), torch.float32)
        ops.split_tf32(kw_bn.view(R, W), a3, 0)
        cos = _new((R, _ceil4(vocab.V)), dev)[:, :vocab.V]
        ops.gemm(a3, vocab.split, out=cos)
        idx = torch.empty(B, K, device=dev, dtype=torch.int64)
        stats = _new((R, 4), dev)
        ops.vq_forward(cos, kw_bn.view(R, W), vocab.norm, vocab.mask_ids, temp, idx, stats)
        return cos, idx, stats

    def forward(self, ws: Workspace, p, audio_feat, kv_len, bn_buffers, vocab: Vocabulary, text: TextTowerPlan, temp: float,
                sot: int, eot: int, training: bool, need_grad: bool, drop=None):
        """-> (text feature fp32 [B, embed], keywords fp32 [B, K, W], cos, idx, stats, saved)."""
        kw_bn, s = self.keywords_forward(ws, p, audio_feat, kv_len, bn_buffers, training, drop)
        cos, idx, stats = self.quantize(ws, vocab, kw_bn, temp)
        B, K, W = kw_bn.shape
        x0 = _new((B, K + 2, W), audio_feat.device)
        keywords = _new((B, K, W), audio_feat.device)
        ops.keyword_embed(vocab.E, text.pos, idx, sot, eot, x0, keywords)
        feat, ts = text.forward(ws, x0, K + 1, save=need_grad)
        s.update(cos=cos, idx=idx, stats=stats, text=ts, temp=temp)
        return feat, keywords, cos, idx, stats, s

    # ------------------------------------------------------------------------------------------------- backward
    def backward(self, ws: Workspace, p, s: dict, dfeat: torch.Tensor, g: Dict[str, torch.Tensor], vocab: Vocabulary,
                 text: TextTowerPlan, need_dfeat: bool = True):
        """dfeat fp32 [B, embed] = d loss / d text feature.  Parameter gradients are written into ``g[name]`` (overwritten);
        returns d audio_feat as a strided fp32 view [B, T, d] (or None)."""
        B, T = s["B"], s["T"]
        d, K, W, hd, heads, dev = self.d, self.K, self.W, self.hd, self.heads, dfeat.device
        Tk, R = T + K, B * K
        M = B * Tk
        # ---- text tower, straight-through estimator, cosine similarity
        dx0 = text.backward(ws, s["text"], dfeat)                      # [B, K+2, W]
        dkeys = ws.view("casc_dkeys", (R, W), torch.float32)
        ops.cast_rows(dx0.view(B, (K + 2) * W)[:, W:(K + 1) * W], dkeys.view(B, K * W), rows=B, cols=K * W, src_ld=(K + 2) * W,
                      dst_ld=K * W)
        V = vocab.V
        gbuf = ws.view("casc_dprob", (R, _ceil4(V)), torch.float32)[:, :V]
        ops.gemm(dkeys, vocab.E, out=gbuf)                              # d loss / d subword_prob = dkeywords @ E^T
        t2 = ws.view("casc_t2", (R,), torch.float32)
        ops.vq_backward(gbuf, s["cos"], s["stats"], s["temp"], t2)      # -> d loss / d cos
        t1g = ws.view("casc_t1", (R, W), torch.float32)
        ops.gemm(gbuf, vocab.unit_t[:, :V], out=t1g)
        dkw_bn = _new((B, K, W), dev)
        ops.cosine_bwd_rows(t1g, t2, s["kw_bn"].view(R, W), s["stats"], dkw_bn.view(R, W))
        # ---- BatchNorm, linear_proj
        dkwp = _new((B, K, W), dev)
        ops.batchnorm_bwd(dkw_bn, s["kwp"], p["bn_layer.bn_layer.weight"], s["mean"], s["rstd"], dkwp, g["bn_layer.bn_layer.weight"],
                          g["bn_layer.bn_layer.bias"])
        dkwp = dkwp.view(R, W)
        wgrad(ws, "cp", dkwp, s["x1"], g["linear_proj.weight"])
        ops.column_sum(dkwp, g["linear_proj.bias"])
        dx1 = _new((R, d), dev)
        dgrad(ws, "cp", dkwp, p["linear_proj.weight"], dx1)
        # ---- attentionBlock_Norm, out-proj (+ [CLS] residual)
        gnw, gnb = g["self_att.attentionBlock_Norm.weight"], g["self_att.attentionBlock_Norm.bias"]
        gnw.zero_()
        gnb.zero_()
        dt1 = _new((R, d), dev)
        ops.layernorm_bwd(dx1, s["t1"], s["st1"], p["self_att.attentionBlock_Norm.weight"], dt1, gnw, gnb)
        wgrad(ws, "co", dt1, s["ctx"].view(R, d), g[A0 + "out_proj.weight"])
        ops.column_sum(dt1, g[A0 + "out_proj.bias"])
        dcls = g["cls"].view(K, d)
        ops.column_sum(dt1, dcls.view(1, K * d), rows=B, cols=K * d, ld=K * d)   # residual path: sum over utterances per keyword row
        dctx = _new((B, K, d), dev)
        dgrad(ws, "co", dt1, p[A0 + "out_proj.weight"], dctx.view(R, d))
        # ---- K-query attention
        w_in = p[A0 + "in_proj_weight"]
        cls = p["cls"].view(K, d)
        kv = s["kv"]
        dkv = ws.view("casc_dkv", (B, Tk, 2 * d), BF)
        g_w, g_b = g[A0 + "in_proj_weight"], g[A0 + "in_proj_bias"]
        dq_part = ws.view("casc_dq_part", (B, K * d), torch.float32)
        ops.mq_attention_bwd(s["q"], kv, 0, d, s["kv_len"], heads, hd, hd ** -0.5, s["probs"], dctx, dkv, dq_part, drop=s.get("drop"))
        dq = ws.view("casc_dq", (K, d), torch.float32)
        ops.column_sum(dq_part, dq.view(1, K * d))
        ops.sgemm(dq.t(), cls.t(), g_w[:d])                            # dWq[o, i] = sum_k dq[k, o] cls[k, i]
        ops.column_sum(dq, g_b[:d])
        ops.sgemm(dq, w_in[:d].t(), dcls, beta=1.0)                    # dcls += dq Wq
        ops.column_sum(dkv.view(M, 2 * d), g_b[d:])
        # ---- K/V projection wgrad (split along the B*(T+K) contraction) and dgrad, bf16 operands on the tensor cores
        NS = 4 if M >= 4096 else 1
        kc = (M + NS * 64 - 1) // (NS * 64) * 64
        ldt = NS * kc
        dkv_t = ws.view(f"casc_dkv_t_{ldt}", (2 * d, ldt), BF, zero=True)
        src_t = ws.view(f"casc_src_t_{ldt}", (d, ldt), BF, zero=True)
        ops.transpose(dkv.view(M, 2 * d), dkv_t[:, :M])
        ops.transpose(s["src"].view(M, d), src_t[:, :M])
        if ldt > M:  # the buffers are shared by every M of this 256-bucket: columns [M, ldt) may hold a LONGER earlier batch
            dkv_t[:, M:].zero_()
            src_t[:, M:].zero_()
        if NS == 1:
            ops.gemm_raw(a=dkv_t, a_inner=M, a_rows=2 * d, a_row_stride=ldt, m_per_batch=2 * d, w=src_t, n=d, k=M, b_row_stride=ldt,
                         out=g_w, out_offset=d * d, ldc=d)
        else:
            parts = ws.view("casc_wgrad_parts", (2 * d, NS * d), torch.float32)
            ops.gemm_raw(a=dkv_t, a_inner=ldt, a_rows=2 * d, a_row_stride=ldt, m_per_batch=2 * d, w=src_t, n=d, k=kc, b_row_stride=ldt,
                         groups=NS, a_group_cols=kc, b_group_stride=kc, out=parts, ldc=NS * d, out_group_cols=d)
            acc = g_w[d:]
            ops.rows_bias_act(parts[:, 0:d], None, parts[:, d:2 * d], NS * d, ops.ACT_NONE, None, acc, rows=2 * d, d=d, x_ld=NS * d, y_ld=d)
            for i in range(2, NS):
                ops.rows_bias_act(acc, None, parts[:, i * d:(i + 1) * d], NS * d, ops.ACT_NONE, None, acc, rows=2 * d, d=d, x_ld=d, y_ld=d)
        wkv_t = ws.view("casc_wkv_t", (d, 2 * d), BF)
        ops.transpose(w_in[d:], wkv_t)
        if need_dfeat:
            dsrc = _new((B, Tk, d), dev)
            ops.gemm(dkv.view(M, 2 * d), wkv_t, out=dsrc.view(M, d))
            ops.column_sum(dsrc, dcls.view(1, K * d), beta=1.0, rows=B, cols=K * d, ld=Tk * d)
            return dsrc[:, K:, :]
        # only the K keyword rows of d src matter (d cls)
        d0 = ws.view("casc_d0", (B, K, d), torch.float32)
        ops.gemm_raw(a=dkv, a_inner=2 * d, a_rows=K, a_row_stride=2 * d, a_batch_stride=Tk * 2 * d, batch=B, m_per_batch=K, w=wkv_t, n=d,
                     k=2 * d, out=d0, ldc=d, out_batch_stride=K * d)
        ops.column_sum(d0, dcls.view(1, K * d), beta=1.0, rows=B, cols=K * d, ld=K * d)
        return None

    # ------------------------------------------------------------------------------------------------- all rows
    def full_forward(self, ws: Workspace, p, audio_feat: torch.Tensor, kv_len: torch.Tensor):
        """KW_CascadedBranch.extract_hidden_states (kwClip.py:829-855): [source, LN(attention + source)] on every row of
        [CLS x K] + audio_feat, fp32 [B, K+T, d]."""
        B, T, d = audio_feat.shape
        src16 = self._build_src(p, audio_feat)
        return self._all_rows(ws, p, src16, kv_len)

    def all_rows(self, ws: Workspace, p, src: torch.Tensor, kv_len: torch.Tensor):
        """MultiheadAttentionAndNorm.forward on an explicit fp32 source [B, L, d] (TransformerModels.py:119-128)."""
        B, L, d = src.shape
        src16 = ws.view("casc_src16", (B, L, d), H)
        ops.cast_rows(src.view(B * L, d), src16.view(B * L, d))
        return self._all_rows(ws, p, src16, kv_len)

    def _all_rows(self, ws: Workspace, p, src16: torch.Tensor, kv_len: torch.Tensor):
        B, Tk, d = src16.shape
        heads, hd, dev = self.heads, self.hd, src16.device
        M = B * Tk
        src16 = src16.view(M, d)
        src32 = _new((B, Tk, d), dev)
        ops.cast_rows(src16, src32.view(M, d))    # the block consumes the fp16-rounded source; expose the same values
        wqkv16 = ws.view("casc_wqkv16", (3 * d, d), H)
        ops.cast_rows(p[A0 + "in_proj_weight"], wqkv16)
        wo16 = ws.view("casc_wo16", (d, d), H)
        ops.cast_rows(p[A0 + "out_proj.weight"], wo16)
        qkv = ws.view("casc_qkv", (M + 8, 3 * d), H, zero=True)[:M].view(B, Tk, 3 * d)   # 8 rows of slack, see _wide_head_attention
        ops.gemm(src16, wqkv16, bias=p[A0 + "in_proj_bias"], out=qkv.view(M, 3 * d))
        ctx = ws.view("casc_ctx16", (B, Tk, d), H)
        if hd in (16, 32, 64, 96, 128):
            ops.attention(qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], ctx, heads, hd ** -0.5, kv_len, False)
        else:
            self._wide_head_attention(ws, qkv, ctx, kv_len)
        y = ws.view("casc_y", (M, d), torch.float32)
        ops.gemm(ctx.view(M, d), wo16, bias=p[A0 + "out_proj.bias"], residual=src32.view(M, d), out=y)
        out = _new((B, Tk, d), dev)
        ops.layernorm(y, p["self_att.attentionBlock_Norm.weight"], p["self_att.attentionBlock_Norm.bias"], y32=out, rows=M, d=d, eps=self.eps)
        return [src32, out]

    def _wide_head_attention(self, ws: Workspace, qkv: torch.Tensor, ctx: torch.Tensor, kv_len: torch.Tensor):
        """Softmax attention for head widths the flash kernels do not cover (one 768 / 1024-wide head): per utterance and
        head, S = Q K^T and P V as tensor-core GEMMs around a row-softmax kernel.  Inference surface only."""
        B, Tk, d3 = qkv.shape
        d, heads, hd = d3 // 3, self.heads, self.hd
        Lp = (Tk + 7) // 8 * 8
        assert qkv.untyped_storage().nbytes() - (qkv.storage_offset() + qkv.numel()) * 2 >= 8 * d3 * 2, "qkv needs 8 rows of slack"
        S = ws.view("casc_S", (B, Tk, Lp), torch.float32)
        P = ws.view("casc_P", (B, Tk, Lp), H)
        vt = ws.view("casc_vt", (hd, Lp), H, zero=True)
        for h in range(heads):
            for b in range(B):
                q, k, v = (qkv[b, :, i * d + h * hd:i * d + (h + 1) * hd] for i in range(3))
                # n is rounded up to a multiple of 8: the extra score columns read the next utterance's key rows (or the
                # buffer's slack) and are never consumed -- the softmax below stops at column Tk
                ops.gemm_raw(a=q, a_inner=hd, a_rows=Tk, a_row_stride=3 * d, m_per_batch=Tk, w=k, n=Lp, k=hd, b_row_stride=3 * d,
                             out=S[b], ldc=Lp, alpha=hd ** -0.5)
            ops.softmax_rows(S.view(B * Tk, Lp), Tk, kv_len, Tk, P.view(B * Tk, Lp), Lp)
            for b in range(B):
                v = qkv[b, :, 2 * d + h * hd:2 * d + (h + 1) * hd]
                ops.transpose(v, vt[:, :Tk])
                ops.gemm_raw(a=P[b], a_inner=Lp, a_rows=Tk, a_row_stride=Lp, m_per_batch=Tk, w=vt, n=hd, k=Lp, b_row_stride=Lp,
                             out=ctx[b, :, h * hd:(h + 1) * hd], ldc=d)
