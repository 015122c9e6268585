"""The trainable parallel branch of SpeechCLIP, forward AND backward, as kernel sequences over the C ABI.

Reference: ``avssl/model/kwClip.py:1076-1108`` (KW_ParallelBranch.forward),
``avssl/module/kw_modules/TransformerModels.py:48-96`` (torch ``nn.TransformerEncoderLayer``, post-LN, erf-GELU,
+ final LayerNorm), ``avssl/util/data_utils.py:4-20`` (key-padding mask).

Only output row 0 of the branch is consumed (``out[:, :1]``, kwClip.py:1103).  For a post-LN layer, row 0 of the
output depends on the other rows only through the keys / values of the attention, and its query is the learned
[CLS] vector — identical for every utterance.  ``cls_forward`` therefore runs: one tensor-core GEMM for K,V of all
rows, a single-query attention per (utterance, head), then a chain of [B, d] row operations in fp32.  The result
equals the full-sequence evaluation within fp tolerance; the B*T-row out-proj / MLP / LayerNorms whose results the
reference discards are never computed.  ``full_forward`` evaluates every row (``extract_hidden_states``).

Parameters stay fp32 (the caller's nn.Parameters).  The K/V projection weight is cast to fp16 per call for the
tensor cores; the gradient operands of its dgrad / wgrad GEMMs are bf16 (fp16 would flush 1e-7-sized gradients).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops
from .engine import H, EncoderLayerPlan, Workspace

BF = torch.bfloat16
L0 = "self_att.model.layers.0."
PARAM_ORDER = ["cls", L0 + "self_attn.in_proj_weight", L0 + "self_attn.in_proj_bias", L0 + "self_attn.out_proj.weight",
               L0 + "self_attn.out_proj.bias", L0 + "linear1.weight", L0 + "linear1.bias", L0 + "linear2.weight",
               L0 + "linear2.bias", L0 + "norm1.weight", L0 + "norm1.bias", L0 + "norm2.weight", L0 + "norm2.bias",
               "self_att.model.norm.weight", "self_att.model.norm.bias", "linear_proj.weight", "linear_proj.bias"]


def _new(shape, dev, dtype=torch.float32, zero=False):
    return (torch.zeros if zero else torch.empty)(shape, device=dev, dtype=dtype)


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


def _ceil4(n: int) -> int:
    return (n + 3) // 4 * 4


def _tc_ok(*ts) -> bool:
    """fp32 2-D operands the tcgen05 GEMM can read as TF32: unit inner stride, 16-byte aligned rows and base."""
    return all(t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in ts)


def linear(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias=None, residual=None):
    """out[M,N] = x[M,K] @ w[N,K]^T (+ bias) (+ residual) on fp32 rows: TF32 tensor cores when the layout allows, else SIMT."""
    N = w.shape[0]
    if N % 8 == 0 and _tc_ok(x, w, out) and (residual is None or _tc_ok(residual)):
        return ops.gemm(x, w, bias=bias, residual=residual, out=out)
    ops.sgemm(x, w, out)
    if bias is not None or residual is not None:
        ops.rows_bias_act(out, bias, residual, residual.stride(0) if residual is not None else 0, ops.ACT_NONE, None, out)
    return out


def transposed(ws: Workspace, name: str, x: torch.Tensor) -> torch.Tensor:
    """fp32 [R, C] -> fp32 view [C, R] of a workspace buffer whose row stride is padded to 16 bytes."""
    R, C = x.shape
    buf = ws.view(name, (C, _ceil4(R)), torch.float32)
    ops.transpose(x, buf[:, :R])
    return buf[:, :R]


def wgrad(ws: Workspace, tag: str, dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor):
    """out[N,K] = dy[M,N]^T @ x[M,K]  (contraction over the M rows)."""
    if out.shape[1] % 8 == 0 and _tc_ok(out):
        return ops.gemm(transposed(ws, tag + "_dyT", dy), transposed(ws, tag + "_xT", x), out=out)
    return ops.sgemm(dy.t(), x.t(), out)


def dgrad(ws: Workspace, tag: str, dy: torch.Tensor, w: torch.Tensor, out: torch.Tensor, residual=None):
    """out[M,K] = dy[M,N] @ w[N,K] (+ residual; residual may be out itself)."""
    if w.shape[1] % 8 == 0 and _tc_ok(dy, out) and (residual is None or _tc_ok(residual)):
        return ops.gemm(dy, transposed(ws, tag + "_wT", w), residual=residual, out=out)
    if residual is not None and residual.data_ptr() != out.data_ptr():
        ops.rows_bias_act(residual, None, None, 0, ops.ACT_NONE, None, out)  # out = residual
    return ops.sgemm(dy, w.t(), out, beta=0.0 if residual is None else 1.0)


# dropout sites of the branch layer (nn.TransformerEncoderLayer in train mode, TransformerModels.py:55-75): the element index of a
# site is the flat index into the tensor named here, restricted to the [CLS] row the head evaluates
SITE_ATTN, SITE_DROPOUT1, SITE_FFN, SITE_DROPOUT2 = 0, 1, 2, 3   # probs [B, heads, T+1] | out-proj [B, d] | GELU(linear1) [B, ffn] | linear2 [B, d]


class ParallelHead:
    """Stateless executor; ``p`` maps the names in PARAM_ORDER to live fp32 CUDA tensors."""

    def __init__(self, d_model: int, nhead: int, eps: float = 1e-5, need_projection: bool = True):
        self.d, self.heads, self.eps = d_model, nhead, eps
        self.hd = d_model // nhead
        self.need_projection = need_projection

    # ------------------------------------------------------------------------------------------------- forward
    def _build_src(self, ws: Workspace, p, audio_feat: torch.Tensor) -> torch.Tensor:
        B, T, d = audio_feat.shape
        Tk = T + 1
        src = _new((B, Tk, d), audio_feat.device, H)  # saved for backward: per call, not workspace
        ops.cast_rows(audio_feat, src.view(B, Tk * d)[:, d:], rows=B, cols=T * d, src_ld=T * d, dst_ld=Tk * d)
        ops.broadcast_row(p["cls"].view(-1), None, src, Tk * d, B, d)
        return src

    def cls_forward(self, ws: Workspace, p: Dict[str, torch.Tensor], audio_feat: torch.Tensor, kv_len: torch.Tensor, drop=None):
        """audio_feat fp32 [B, T, d] contiguous; kv_len int32 [B] = audio_len + 1.  -> (out fp32 [B, out_dim], saved).
        drop = (p, rng_state) in train mode: the layer's four dropouts (attention weights, after out-proj, after the
        activation, after linear2) act on the [CLS] row; the masks are functions of rng_state (scb_dropout_mask), which the
        backward pass reads again from ``saved``."""
        B, T, d = audio_feat.shape
        dp, rng = (float(drop[0]), drop[1].clone()) if drop is not None and drop[0] > 0 else (0.0, None)
        site = (lambda k: (dp, rng, k)) if rng is not None else (lambda k: None)
        Tk, hd, heads, dev = T + 1, self.hd, self.heads, audio_feat.device
        M = B * Tk
        w_in, b_in = p[L0 + "self_attn.in_proj_weight"], p[L0 + "self_attn.in_proj_bias"]
        src = self._build_src(ws, p, audio_feat)
        wkv16 = ws.view("head_wkv16", (2 * d, d), H)
        ops.cast_rows(w_in[d:], wkv16)
        kv = _new((B, Tk, 2 * d), dev, H)
        ops.gemm(src.view(M, d), wkv16, bias=b_in[d:], out=kv.view(M, 2 * d))
        cls = p["cls"].view(1, d)
        q = _new((1, d), dev)
        ops.sgemm(cls, w_in[:d], q)
        ops.rows_bias_act(q, b_in[:d], None, 0, ops.ACT_NONE, None, q)
        probs = _new((B, heads, Tk), dev)
        ctx = _new((B, d), dev)
        ops.cls_attention_fwd(q, kv, 0, d, kv_len, heads, hd, hd ** -0.5, probs, ctx, drop=site(SITE_ATTN))
        t1 = _new((B, d), dev)
        linear(ctx, p[L0 + "self_attn.out_proj.weight"], t1, bias=p[L0 + "self_attn.out_proj.bias"])
        if rng is not None:
            ops.dropout_rows(t1, t1, site(SITE_DROPOUT1))
        ops.rows_bias_act(t1, None, cls, 0, ops.ACT_NONE, None, t1)  # + [CLS] residual (one row, broadcast)
        x1, st1 = _new((B, d), dev), _new((B, 2), dev)
        ops.layernorm(t1, p[L0 + "norm1.weight"], p[L0 + "norm1.bias"], y32=x1, stats=st1, eps=self.eps)
        ffn = p[L0 + "linear1.weight"].shape[0]
        h_pre, h = _new((B, ffn), dev), _new((B, ffn), dev)
        linear(x1, p[L0 + "linear1.weight"], h_pre, bias=p[L0 + "linear1.bias"])
        ops.rows_bias_act(h_pre, None, None, 0, ops.ACT_GELU, None, h)
        t2 = _new((B, d), dev)
        if rng is not None:
            ops.dropout_rows(h, h, site(SITE_FFN))
            linear(h, p[L0 + "linear2.weight"], t2, bias=p[L0 + "linear2.bias"])
            ops.dropout_rows(t2, t2, site(SITE_DROPOUT2))
            ops.rows_bias_act(t2, None, x1, x1.stride(0), ops.ACT_NONE, None, t2)
        else:
            linear(h, p[L0 + "linear2.weight"], t2, bias=p[L0 + "linear2.bias"], residual=x1)
        x2, st2 = _new((B, d), dev), _new((B, 2), dev)
        ops.layernorm(t2, p[L0 + "norm2.weight"], p[L0 + "norm2.bias"], y32=x2, stats=st2, eps=self.eps)
        x3, st3 = _new((B, d), dev), _new((B, 2), dev)
        ops.layernorm(x2, p["self_att.model.norm.weight"], p["self_att.model.norm.bias"], y32=x3, stats=st3, eps=1e-5)
        if self.need_projection:
            out = _new((B, p["linear_proj.weight"].shape[0]), dev)
            linear(x3, p["linear_proj.weight"], out, bias=p["linear_proj.bias"])
        else:
            out = x3
        saved = dict(B=B, T=T, src=src, kv=kv, q=q, probs=probs, ctx=ctx, t1=t1, x1=x1, st1=st1, h_pre=h_pre, h=h, t2=t2, x2=x2,
                     st2=st2, st3=st3, x3=x3, kv_len=kv_len, drop=(dp, rng))
        return out, saved

    # ------------------------------------------------------------------------------------------------- backward
    def cls_backward(self, ws: Workspace, p: Dict[str, torch.Tensor], s: dict, dout: torch.Tensor, g: Dict[str, torch.Tensor],
                     need_dfeat: bool = True):
        """dout fp32 [B, out_dim].  Writes parameter gradients into the fp32 tensors ``g[name]`` (overwritten) and returns
        d audio_feat as a strided fp32 view [B, T, d] (rows 1.. of the source gradient)."""
        B, T = s["B"], s["T"]
        d, hd, heads, dev = self.d, self.hd, self.heads, dout.device
        Tk = T + 1
        M = B * Tk
        w_in = p[L0 + "self_attn.in_proj_weight"]
        cls = p["cls"].view(1, d)
        dcls = g["cls"].view(1, d)
        dp, rng = s.get("drop", (0.0, None))
        site = (lambda k: (dp, rng, k)) if rng is not None else (lambda k: None)
        if self.need_projection:
            wp = p["linear_proj.weight"]
            wgrad(ws, "hp", dout, s["x3"], g["linear_proj.weight"])
            ops.column_sum(dout, g["linear_proj.bias"])
            dx3 = _new((B, d), dev)
            dgrad(ws, "hp", dout, wp, dx3)
        else:
            dx3 = dout.contiguous()
        # final norm, norm2
        for name in ("self_att.model.norm.", L0 + "norm2.", L0 + "norm1."):
            g[name + "weight"].zero_()
            g[name + "bias"].zero_()
        dx2 = _new((B, d), dev)
        ops.layernorm_bwd(dx3, s["x2"], s["st3"], p["self_att.model.norm.weight"], dx2, g["self_att.model.norm.weight"],
                          g["self_att.model.norm.bias"])
        dt2 = _new((B, d), dev)
        ops.layernorm_bwd(dx2, s["t2"], s["st2"], p[L0 + "norm2.weight"], dt2, g[L0 + "norm2.weight"], g[L0 + "norm2.bias"])
        # MLP (train mode: the gradient of the linear2 path passes dropout2's mask, that of the activation dropout's; the
        # residual path of dt2 does not)
        dy2 = ops.dropout_rows(dt2, _new((B, d), dev), site(SITE_DROPOUT2)) if rng is not None else dt2
        wgrad(ws, "h2", dy2, s["h"], g[L0 + "linear2.weight"])
        ops.column_sum(dy2, g[L0 + "linear2.bias"])
        dh = _new(s["h"].shape, dev)
        dgrad(ws, "h2", dy2, p[L0 + "linear2.weight"], dh)
        if rng is not None:
            ops.dropout_rows(dh, dh, site(SITE_FFN))
        ops.gelu_bwd(dh, s["h_pre"], dh)
        wgrad(ws, "h1", dh, s["x1"], g[L0 + "linear1.weight"])
        ops.column_sum(dh, g[L0 + "linear1.bias"])
        dgrad(ws, "h1", dh, p[L0 + "linear1.weight"], dt2, residual=dt2)  # dx1 = dt2 (residual path) + dh W1
        dt1 = _new((B, d), dev)
        ops.layernorm_bwd(dt2, s["t1"], s["st1"], p[L0 + "norm1.weight"], dt1, g[L0 + "norm1.weight"], g[L0 + "norm1.bias"])
        # attention out-proj (+ residual = [CLS]; dropout1 sits on the out-proj path only)
        dy1 = ops.dropout_rows(dt1, _new((B, d), dev), site(SITE_DROPOUT1)) if rng is not None else dt1
        wgrad(ws, "ho", dy1, s["ctx"], g[L0 + "self_attn.out_proj.weight"])
        ops.column_sum(dy1, g[L0 + "self_attn.out_proj.bias"])
        ops.column_sum(dt1, dcls)
        dctx = _new((B, d), dev)
        dgrad(ws, "ho", dy1, p[L0 + "self_attn.out_proj.weight"], dctx)
        # single-query attention
        kv = s["kv"]
        dkv = ws.view("head_dkv", (B, Tk, 2 * d), BF)
        g_w, g_b = g[L0 + "self_attn.in_proj_weight"], g[L0 + "self_attn.in_proj_bias"]
        dq = g_b[:d]
        dq.zero_()
        ops.cls_attention_bwd(s["q"], kv, 0, d, s["kv_len"], heads, hd, hd ** -0.5, s["probs"], dctx, dkv, dq, drop=site(SITE_ATTN))
        ops.sgemm(dq.view(d, 1), cls.view(d, 1), g_w[:d])            # dWq = dq (x) cls
        ops.sgemm(dq.view(1, d), w_in[:d].t(), dcls, beta=1.0)        # dcls += Wq^T dq
        ops.column_sum(dkv.view(M, 2 * d), g_b[d:])
        # K/V projection: wgrad dWkv = dKV^T src, dgrad dsrc = dKV Wkv  (bf16 operands on the tensor cores).
        # The wgrad contracts over all M = B*(T+1) rows into only 12 x 3 output tiles, so it is split NS ways along M
        # (the GEMM's group dimension walks the M slices of both operands) into NS partial [2d, d] results that are summed.
        NS = 4 if M >= 4096 else 1
        kc = (M + NS * 64 - 1) // (NS * 64) * 64   # rows per split, multiple of the 64-wide k-block
        ldt = NS * kc                               # padded row length: every split stays inside its own row
        dkv_t = ws.view(f"head_dkv_t_{ldt}", (2 * d, ldt), BF, zero=True)
        src_t = ws.view(f"head_src_t_{ldt}", (d, ldt), BF, zero=True)
        ops.transpose(dkv.view(M, 2 * d), dkv_t[:, :M])
        ops.transpose(s["src"].view(M, d), src_t[:, :M])
        if ldt > M:  # the buffers are shared by every M of this 256-bucket: columns [M, ldt) may hold a LONGER earlier batch
            dkv_t[:, M:].zero_()
            src_t[:, M:].zero_()
        if NS == 1:
            ops.gemm_raw(a=dkv_t, a_inner=M, a_rows=2 * d, a_row_stride=ldt, m_per_batch=2 * d, w=src_t, n=d, k=M, b_row_stride=ldt,
                         out=g_w, out_offset=d * d, ldc=d)
        else:
            parts = ws.view("head_wgrad_parts", (2 * d, NS * d), torch.float32)
            ops.gemm_raw(a=dkv_t, a_inner=ldt, a_rows=2 * d, a_row_stride=ldt, m_per_batch=2 * d, w=src_t, n=d, k=kc, b_row_stride=ldt,
                         groups=NS, a_group_cols=kc, b_group_stride=kc, out=parts, ldc=NS * d, out_group_cols=d)
            acc = g_w[d:]
            ops.rows_bias_act(parts[:, 0:d], None, parts[:, d:2 * d], NS * d, ops.ACT_NONE, None, acc, rows=2 * d, d=d, x_ld=NS * d, y_ld=d)
            for i in range(2, NS):
                ops.rows_bias_act(acc, None, parts[:, i * d:(i + 1) * d], NS * d, ops.ACT_NONE, None, acc, rows=2 * d, d=d, x_ld=d, y_ld=d)
        dsrc = None
        if need_dfeat:
            wkv_t = ws.view("head_wkv_t", (d, 2 * d), BF)
            ops.transpose(w_in[d:], wkv_t)
            dsrc = _new((B, Tk, d), dev)
            ops.gemm(dkv.view(M, 2 * d), wkv_t, out=dsrc.view(M, d))
            ops.column_sum(dsrc, dcls, beta=1.0, rows=B, cols=d, ld=Tk * d)
            return dsrc[:, 1:, :]
        # without a consumer for d audio_feat only row 0 of dsrc matters (d cls)
        wkv_t = ws.view("head_wkv_t", (d, 2 * d), BF)
        ops.transpose(w_in[d:], wkv_t)
        d0 = _new((B, d), dev)
        ops.gemm_raw(a=dkv, a_inner=2 * d, a_rows=B, a_row_stride=Tk * 2 * d, m_per_batch=B, w=wkv_t, n=d, k=2 * d, out=d0, ldc=d)
        ops.column_sum(d0, dcls, beta=1.0)
        return None

    # ------------------------------------------------------------------------------------------------- all rows
    def full_forward(self, ws: Workspace, p: Dict[str, torch.Tensor], audio_feat: torch.Tensor, kv_len: torch.Tensor
                     ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """Every row of the layer on [CLS] + audio_feat (KW_ParallelBranch.extract_hidden_states, kwClip.py:1049-1074)."""
        B, T, d = audio_feat.shape
        src16 = self._build_src(ws, p, audio_feat)
        return self._full(ws, p, src16, kv_len)

    def full_forward_src(self, ws: Workspace, p: Dict[str, torch.Tensor], src: torch.Tensor, kv_len: torch.Tensor):
        """Every row of the layer on an explicit fp32 source [B, L, d] (TransformerEncoder.forward, TransformerModels.py:77-96)."""
        B, L, d = src.shape
        src16 = ws.view("head_src", (B, L, d), H)
        ops.cast_rows(src.view(B * L, d), src16.view(B * L, d))
        return self._full(ws, p, src16, kv_len)

    def _full(self, ws: Workspace, p, src16: torch.Tensor, kv_len: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """-> (final-norm output fp32 [B, L, d], [layer input, layer output] fp32)  (TransformerModels.py:16-45)."""
        B, Tk, d = src16.shape
        dev = src16.device
        M = B * Tk
        src16 = src16.view(M, d)
        src32 = _new((B, Tk, d), dev)
        ops.cast_rows(src16, src32.view(M, d))  # the layer consumes the fp16-rounded source; expose the same values
        layer = EncoderLayerPlan(dev, wqkv=p[L0 + "self_attn.in_proj_weight"], bqkv=p[L0 + "self_attn.in_proj_bias"],
                                 wo=p[L0 + "self_attn.out_proj.weight"], bo=p[L0 + "self_attn.out_proj.bias"],
                                 ln1=(p[L0 + "norm1.weight"], p[L0 + "norm1.bias"]), w1=p[L0 + "linear1.weight"],
                                 b1=p[L0 + "linear1.bias"], w2=p[L0 + "linear2.weight"], b2=p[L0 + "linear2.bias"],
                                 ln2=(p[L0 + "norm2.weight"], p[L0 + "norm2.bias"]), heads=self.heads, pre_ln=False,
                                 act=ops.ACT_GELU, eps=self.eps)
        out1 = _new((B, Tk, d), dev)
        layer.forward(ws, src32.view(M, d), src16, out1.view(M, d), B, Tk, kv_len, want_x16=False, tag="head_")
        final = _new((B, Tk, d), dev)
        ops.layernorm(out1, p["self_att.model.norm.weight"], p["self_att.model.norm.bias"], y32=final, rows=M, d=d, eps=1e-5)
        return final, [src32, out1]
