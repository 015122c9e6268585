"""Thin tensor-level wrappers over the C ABI: take torch CUDA tensors, pass raw pointers + sizes.

torch is used for device memory and the current stream only; every arithmetic op below runs in
libspeechclip_b200.so (include/speechclip_b200.h).  There is no CPU path: CPU tensors assert.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import lib as _l

_DT = {torch.float32: _l.F32, torch.float16: _l.F16, torch.bfloat16: _l.BF16}
F32, F16, BF16 = _l.F32, _l.F16, _l.BF16
ACT_NONE, ACT_GELU, ACT_QUICK_GELU = _l.ACT_NONE, _l.ACT_GELU, _l.ACT_QUICK_GELU


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda, "CUDA tensor required (there is no CPU path)"
    return ctypes.c_void_p(t.data_ptr())


# bench.py's live per-entry-point timing: when PROFILE is a list every C-ABI call is bracketed by CUDA events on the
# launching stream and (name, algorithmic flops, ev0, ev1) is appended.  None (the default) adds no work.
PROFILE = None
_FLOPS = 0.0
_SHAPE = ""


def _call(name: str, *args):
    global _FLOPS, _SHAPE
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(_l.load(), name)(*args, _stream())
    if prof is not None:
        e1.record()
        prof.append((name, _FLOPS, e0, e1, _SHAPE))
        _FLOPS, _SHAPE = 0.0, ""
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {_l.load().scb_last_error().decode()}")


# ------------------------------------------------------------------------------------------------ tensor-core GEMM
def gemm_raw(*, a: torch.Tensor, a_inner: int, a_rows: int, a_row_stride: int, a_batch_stride: int = 0, batch: int = 1,
             m_per_batch: int, w: torch.Tensor, n: int, k: int, b_row_stride: Optional[int] = None, groups: int = 1,
             b_group_stride: int = 0, kb_per_tap: Optional[int] = None, tap_row_shift: int = 0, a_col0: int = 0,
             a_group_cols: int = 0, out: torch.Tensor, ldc: int, out_batch_stride: int = 0, out_group_cols: int = 0,
             out2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
             residual: Optional[torch.Tensor] = None, residual_ld: int = 0, residual_batch_stride: int = 0,
             act: int = ACT_NONE, alpha: float = 1.0, a_offset: int = 0, out_offset: int = 0, residual_offset: int = 0,
             algo_k: Optional[int] = None, scratch: Optional[torch.Tensor] = None):
    """Full operand model of scb_gemm (plain / strided-conv / grouped tap walk); offsets are in elements.
    scratch: `gemm_workspace_bytes()` of zero-initialised device memory private to the launching stream (stream-K tail).
    algo_k: the ALGORITHMIC contraction length when k carries zero padding (pos-conv groups padded 48 -> 64 channels)."""
    global _FLOPS, _SHAPE
    if PROFILE is not None:
        _FLOPS = 2.0 * batch * groups * m_per_batch * n * (algo_k if algo_k is not None else k)
        _SHAPE = f"b{batch} g{groups} m{m_per_batch} n{n} k{k} {str(a.dtype)[6:]}->{str(out.dtype)[6:]} act{act}{' bias' if bias is not None else ''}{' res' if residual is not None else ''}"
    assert a.dtype in (torch.float16, torch.bfloat16, torch.float32) and w.dtype == a.dtype
    g = _l.GemmArgs()
    g.a = a.data_ptr() + a_offset * a.element_size()
    g.a_inner, g.a_rows, g.a_row_stride, g.a_batch_stride = a_inner, a_rows, a_row_stride, a_batch_stride
    g.batch, g.m_per_batch = batch, m_per_batch
    bk = 128 // a.element_size()
    g.kb_per_tap = kb_per_tap if kb_per_tap is not None else (k + bk - 1) // bk
    g.tap_row_shift, g.a_col0, g.a_group_cols = tap_row_shift, a_col0, a_group_cols
    g.b = w.data_ptr()
    g.b_row_stride = b_row_stride if b_row_stride is not None else k
    g.b_group_stride, g.n, g.k, g.groups = b_group_stride, n, k, groups
    g.out = out.data_ptr() + out_offset * out.element_size()
    g.out_dtype, g.out_group_cols, g.ldc, g.out_batch_stride = _DT[out.dtype], out_group_cols, ldc, out_batch_stride
    if out2 is not None:
        g.out2, g.out2_dtype = out2.data_ptr() + out_offset * out2.element_size(), _DT[out2.dtype]
    g.ab_format = _DT[a.dtype]
    if bias is not None:
        assert bias.dtype == torch.float32
        g.bias = bias.data_ptr()
    if residual is not None:
        g.residual = residual.data_ptr() + residual_offset * residual.element_size()
        g.residual_dtype = _DT[residual.dtype]
        g.residual_ld, g.residual_batch_stride = residual_ld, residual_batch_stride
    g.act, g.alpha = act, alpha
    if scratch is not None:
        g.workspace, g.workspace_bytes = scratch.data_ptr(), scratch.numel() * scratch.element_size()
    _call("scb_gemm", ctypes.byref(g))
    return out


_GEMM_WS_BYTES = None


def gemm_workspace_bytes() -> int:
    global _GEMM_WS_BYTES
    if _GEMM_WS_BYTES is None:
        _GEMM_WS_BYTES = int(_l.load().scb_gemm_workspace_bytes())
    return _GEMM_WS_BYTES


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.float16, out2: Optional[torch.Tensor] = None, alpha: float = 1.0,
         scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = act(alpha * a[M,K] @ w[N,K]^T + bias) + residual   (a, w 16-bit; fp32 accumulate on tcgen05)."""
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    if out2 is not None:
        assert out2.shape == out.shape and out2.stride(0) == out.stride(0)
    res_ld = 0
    if residual is not None:
        assert residual.shape == out.shape and residual.stride(1) == 1
        res_ld = residual.stride(0)
    return gemm_raw(a=a, a_inner=K, a_rows=M, a_row_stride=a.stride(0), m_per_batch=M, w=w, n=N, k=K, b_row_stride=w.stride(0),
                    out=out, ldc=out.stride(0), out2=out2, bias=bias, residual=residual, residual_ld=res_ld, act=act, alpha=alpha,
                    scratch=scratch)


def sgemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, alpha: float = 1.0, beta: float = 0.0):
    """c[M,N] = alpha * sum_k a[m,k] b[n,k] + beta * c  — a, b arbitrary-strided 2-D fp32 views (pass .t() for transposes)."""
    assert a.dtype == b.dtype == c.dtype == torch.float32 and a.dim() == b.dim() == c.dim() == 2
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and c.shape == (M, N) and c.stride(1) == 1
    _call("scb_sgemm", _p(a), a.stride(0), a.stride(1), _p(b), b.stride(0), b.stride(1), _p(c), c.stride(0), M, N, K, alpha, beta)
    return c


# ------------------------------------------------------------------------------------------------ attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, heads: int, scale: float,
              kv_len: Optional[torch.Tensor] = None, causal: bool = False):
    """q/k/v/out: [B, T, heads*hd] 16-bit views (last dim contiguous; typically slices of one fused QKV buffer)."""
    B, Tq, D = q.shape
    Tk = k.shape[1]
    hd = D // heads
    for t in (q, k, v, out):
        assert t.stride(2) == 1 and t.dtype == q.dtype
    if kv_len is not None:
        assert kv_len.dtype == torch.int32
    _call("scb_attention_fwd", _p(q), _p(k), _p(v), _p(out), _DT[q.dtype], q.stride(1), k.stride(1), v.stride(1), out.stride(1),
          q.stride(0), k.stride(0), v.stride(0), out.stride(0), _p(kv_len), B, heads, hd, Tq, Tk, scale, int(causal))
    return out


def cls_attention_fwd(q, kv, k_off, v_off, kv_len, heads, hd, scale, probs, ctx32, drop=None):
    """drop = (p, rng_state int64[2] on the device, site) for train-mode attention dropout, or None."""
    B, Tk = kv.shape[0], kv.shape[1]
    p, state, site = drop if drop is not None else (0.0, None, 0)
    _call("scb_cls_attention_fwd", _p(q), _p(kv), _DT[kv.dtype], kv.stride(1), kv.stride(0), k_off, v_off, _p(kv_len), B, heads, hd,
          Tk, scale, _p(probs), _p(ctx32), None, 0, p, _p(state), site)


def cls_attention_bwd(q, kv, k_off, v_off, kv_len, heads, hd, scale, probs, dctx, dkv, dq, drop=None):
    B, Tk = kv.shape[0], kv.shape[1]
    assert dkv.shape == kv.shape and dkv.stride() == kv.stride()
    p, state, site = drop if drop is not None else (0.0, None, 0)
    _call("scb_cls_attention_bwd", _p(q), _p(kv), _DT[kv.dtype], kv.stride(1), kv.stride(0), k_off, v_off, _p(kv_len), B, heads, hd,
          Tk, scale, _p(probs), _p(dctx), _p(dkv), _DT[dkv.dtype], _p(dq), p, _p(state), site)


# ------------------------------------------------------------------------------------------------ dropout (trainable branch)
def rng_advance(state):
    assert state.dtype == torch.int64 and state.numel() == 2
    _call("scb_rng_advance", _p(state))


def dropout_mask(state, site, p, n):
    """The multiplicative mask (0 or 1/(1-p), fp32 [n]) of dropout site ``site`` at the RNG state ``state``."""
    mask = torch.empty(n, device=state.device, dtype=torch.float32)
    _call("scb_dropout_mask", _p(state), site, p, _p(mask), n)
    return mask


def dropout_rows(x, y, drop):
    """y = x * mask(drop) elementwise over a contiguous fp32 tensor (in place allowed); drop = (p, rng_state, site)."""
    p, state, site = drop
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == y.dtype == torch.float32
    _call("scb_dropout_rows", _p(x), _p(y), x.numel(), p, _p(state), site)
    return y


# ------------------------------------------------------------------------------------------------ front end
def frame_lengths(wav_len, batch, tw_out, max_audio_len, n_frames, rate, u, crop_off, crop_len, valid_frames, feat_len, feat_len64):
    _call("scb_frame_lengths", _p(wav_len), batch, tw_out, max_audio_len, n_frames, rate, _p(u), _p(crop_off), _p(crop_len),
          _p(valid_frames), _p(feat_len), _p(feat_len64))


def lengths_to_i32(src, add, clamp_max, out):
    assert src.dtype == torch.int64 and out.dtype == torch.int32
    _call("scb_lengths_to_i32", _p(src), src.numel(), add, clamp_max, _p(out))
    return out


def wav_prepare(wav, crop_off, crop_len, tw_out, normalize, stats_scratch, out):
    assert wav.dtype == torch.float32 and wav.stride(1) == 1 and out.stride(1) == 1
    _call("scb_wav_prepare", _p(wav), wav.stride(0), wav.shape[0], _p(crop_off), _p(crop_len), tw_out, int(normalize),
          _p(stats_scratch), _p(out), out.stride(0))


def conv0_scratch_bytes(batch: int) -> int:
    return int(_l.load().scb_conv0_scratch_bytes(batch))


def conv0_groupnorm_gelu(wav, n_samples, w, conv_bias, gamma, beta, eps, out, out_batch_stride, scratch):
    _call("scb_conv0_groupnorm_gelu", _p(wav), wav.stride(0), wav.shape[0], n_samples, _p(w), _p(conv_bias), _p(gamma), _p(beta), eps,
          _p(out), _DT[out.dtype], out_batch_stride, _p(scratch), scratch.numel() * scratch.element_size())


def conv0_layernorm_gelu(wav, n_samples, w, conv_bias, gamma, beta, eps, out, out_batch_stride, scratch):
    _call("scb_conv0_layernorm_gelu", _p(wav), wav.stride(0), wav.shape[0], n_samples, _p(w), _p(conv_bias), _p(gamma), _p(beta), eps,
          _p(out), _DT[out.dtype], out_batch_stride, _p(scratch), scratch.numel() * scratch.element_size())


def posconv_pack(x, valid_frames, xpad, batch, T, D, groups, pad_left, rows_pad):
    _call("scb_posconv_pack", _p(x), _p(valid_frames), _p(xpad), _DT[xpad.dtype], batch, T, D, groups, pad_left, rows_pad)


def patchify(img, out, P, ldk):
    B, C, H, W = img.shape
    assert img.is_contiguous() and img.dtype == torch.float32
    _call("scb_patchify", _p(img), _p(out), _DT[out.dtype], B, C, H, W, P, ldk)


def broadcast_row(a, a2, out, out_stride, nb, d, out_offset=0):
    o = ctypes.c_void_p(out.data_ptr() + out_offset * out.element_size())
    _call("scb_broadcast_row", _p(a), _p(a2), o, _DT[out.dtype], out_stride, nb, d)


def cast_rows(src, dst, rows=None, cols=None, src_ld=None, dst_ld=None):
    """dst[r, c] = src[r, c] with dtype conversion; 2-D views with contiguous last dim (or explicit rows/cols/ld)."""
    if rows is None:
        rows, cols = src.shape
        src_ld, dst_ld = src.stride(0), dst.stride(0)
    _call("scb_cast_rows", _p(src), _DT[src.dtype], src_ld, _p(dst), _DT[dst.dtype], dst_ld, rows, cols)
    return dst


def transpose(src, dst):
    """dst[c, r] = src[r, c] (2-D, last dim contiguous), converting dtype."""
    rows, cols = src.shape
    assert dst.shape == (cols, rows) and src.stride(1) == 1 and dst.stride(1) == 1
    _call("scb_transpose", _p(src), _DT[src.dtype], src.stride(0), _p(dst), _DT[dst.dtype], dst.stride(0), rows, cols)
    return dst


# ------------------------------------------------------------------------------------------------ row kernels
def layernorm(x, gamma, beta, *, y32=None, y16=None, stats=None, rows=None, d=None, x_ld=None, y_ld=None, eps=1e-5, act=ACT_NONE):
    if rows is None:
        d = x.shape[-1]
        rows = x.numel() // d
    x_ld = d if x_ld is None else x_ld
    y_ld = d if y_ld is None else y_ld
    _call("scb_layernorm_fwd", _p(x), _DT[x.dtype], _p(gamma), _p(beta), _p(y32), _p(y16), _DT[y16.dtype] if y16 is not None else 0,
          _p(stats), rows, d, x_ld, y_ld, eps, act)


def layernorm_bwd(dy, x, stats, gamma, dx, dgamma, dbeta):
    d = x.shape[-1]
    _call("scb_layernorm_bwd", _p(dy), _p(x), _p(stats), _p(gamma), _p(dx), _p(dgamma), _p(dbeta), x.numel() // d, d)


def l2norm(x, y, norms):
    _call("scb_l2norm_fwd", _p(x), _p(y), _p(norms), x.shape[0], x.shape[1])


def l2norm_bwd(dy, y, norms, dx):
    _call("scb_l2norm_bwd", _p(dy), _p(y), _p(norms), _p(dx), y.shape[0], y.shape[1])


def weighted_sum(h, w_logits, normalize, *, out32=None, out16=None, rows_per_batch=0, out16_batch_stride=0, out16_row0=0):
    """h: [L, rows, d] fp32 or fp16 (contiguous)."""
    L, rows, d = h.shape
    _call("scb_weighted_sum_fwd", _p(h), _DT[h.dtype], h.stride(0), _p(w_logits), L, int(normalize), _p(out32), _p(out16),
          _DT[out16.dtype] if out16 is not None else 0, rows, d, rows_per_batch, out16_batch_stride, out16_row0)


def weighted_sum_bwd(h, w_logits, normalize, dout, rows_per_batch, dout_batch_stride, dout_row0, scratch_L, grad_logits, grad_scale=1.0):
    L, rows, d = h.shape
    _call("scb_weighted_sum_bwd", _p(h), _DT[h.dtype], h.stride(0), _p(w_logits), L, int(normalize), _p(dout), rows, d, rows_per_batch,
          dout_batch_stride, dout_row0, _p(scratch_L), _p(grad_logits), grad_scale)


def rows_bias_act(x, bias, res, res_ld, act, pre, y, rows=None, d=None, x_ld=None, y_ld=None):
    if rows is None:
        rows, d = x.shape
        x_ld, y_ld = x.stride(0), y.stride(0)
    _call("scb_rows_bias_act", _p(x), x_ld, _p(bias), _p(res), res_ld, act, _p(pre), _p(y), y_ld, rows, d)


def gelu_bwd(dy, pre, dx):
    _call("scb_gelu_bwd", _p(dy), _p(pre), _p(dx), pre.numel())


def column_sum(x, out, beta=0.0, rows=None, cols=None, ld=None):
    if rows is None:
        rows, cols = x.shape
        ld = x.stride(0)
    _call("scb_column_sum", _p(x), _DT[x.dtype], ld, rows, cols, _p(out), beta)


# ------------------------------------------------------------------------------------------------ loss / optimiser / retrieval
def infonce_scratch_bytes(B: int) -> int:
    return int(_l.load().scb_infonce_scratch_bytes(B))


def infonce(a, b, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a, scratch, *, phase=3, loss=None, logits_out=None, upstream=1.0,
            upstream_dev=None, dA=None, dB=None, dlog_mult=None):
    B, D = a.shape
    assert a.dtype == b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    if ids is not None:
        assert ids.dtype == torch.int64 and ids.is_contiguous()
    _call("scb_infonce", _p(a), _p(b), _p(ids), B, D, _p(log_mult), fixed_mult, margin, int(dcl), int(a2b), int(b2a), phase, _p(loss),
          _p(logits_out), upstream, _p(upstream_dev), _p(dA), _p(dB), _p(dlog_mult), _p(scratch),
          scratch.numel() * scratch.element_size())


def adam_step(p, g, m, v, sumsq, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, step, p_f16=None, p_bf16=None):
    _call("scb_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(sumsq), grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay,
          step, _p(p_f16), _p(p_bf16))


def retrieval_rank(score, cand_ids, answers, rank, top1):
    rows, cols = score.shape
    assert score.dtype == torch.float32 and score.stride(1) == 1
    _call("scb_retrieval_rank", _p(score), score.stride(0), rows, cols, _p(cand_ids), _p(answers), _p(rank), _p(top1))


# ------------------------------------------------------------------------------------------------ cascaded branch
def mq_attention_fwd(q, kv, k_off, v_off, kv_len, heads, hd, scale, probs, ctx32, drop=None):
    """q fp32 [NQ, heads*hd]; kv 16-bit [B, Tk, ld]; probs fp32 [B, heads, NQ, Tk]; ctx32 fp32 [B, NQ, heads*hd];
    drop = (p, rng_state, site) for train-mode attention dropout, or None."""
    B, Tk = kv.shape[0], kv.shape[1]
    p, state, site = drop if drop is not None else (0.0, None, 0)
    _call("scb_mq_attention_fwd", _p(q), _p(kv), _DT[kv.dtype], kv.stride(1), kv.stride(0), k_off, v_off, _p(kv_len), B, heads, hd,
          q.shape[0], Tk, scale, _p(probs), _p(ctx32), p, _p(state), site)


def mq_attention_bwd(q, kv, k_off, v_off, kv_len, heads, hd, scale, probs, dctx, dkv, dq_part, drop=None):
    """dq_part fp32 [B, NQ * heads*hd]: per-utterance contributions to dq (sum over dim 0 with column_sum)."""
    B, Tk = kv.shape[0], kv.shape[1]
    assert dkv.shape == kv.shape and dkv.stride() == kv.stride() and dq_part.numel() == B * q.numel()
    p, state, site = drop if drop is not None else (0.0, None, 0)
    _call("scb_mq_attention_bwd", _p(q), _p(kv), _DT[kv.dtype], kv.stride(1), kv.stride(0), k_off, v_off, _p(kv_len), B, heads, hd,
          q.shape[0], Tk, scale, _p(probs), _p(dctx), _p(dkv), _DT[dkv.dtype], _p(dq_part), p, _p(state), site)


def batchnorm_fwd(x, y, gamma, beta, running_mean, running_var, save_mean, save_rstd, eps, momentum, training):
    B, NK, D = x.shape
    _call("scb_batchnorm_fwd", _p(x), _p(y), _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(save_mean), _p(save_rstd),
          B, NK, D, eps, momentum, int(training))


def batchnorm_bwd(dy, x, gamma, save_mean, save_rstd, dx, dgamma, dbeta):
    B, NK, D = x.shape
    _call("scb_batchnorm_bwd", _p(dy), _p(x), _p(gamma), _p(save_mean), _p(save_rstd), _p(dx), _p(dgamma), _p(dbeta), B, NK, D)


def split_tf32(src, dst, role):
    rows, cols = src.shape
    assert dst.shape == (rows, 3 * cols) and dst.is_contiguous() and src.stride(1) == 1
    _call("scb_split_tf32", _p(src), src.stride(0), _p(dst), rows, cols, role)
    return dst


def vq_forward(dots, kw, emb_norm, mask_ids, temp, idx, stats):
    R, V = dots.shape
    _call("scb_vq_forward", _p(dots), _p(kw), _p(emb_norm), R, V, kw.shape[1] if kw is not None else 0, dots.stride(0), _p(mask_ids),
          0 if mask_ids is None else mask_ids.numel(), temp, _p(idx), _p(stats))


def vq_backward(g, cos, stats, temp, t2):
    R, V = g.shape
    assert g.stride(0) == cos.stride(0)
    _call("scb_vq_backward", _p(g), _p(cos), R, V, g.stride(0), _p(stats), temp, _p(t2))


def cosine_bwd_rows(t1, t2, kw, stats, dkw):
    R, D = kw.shape
    _call("scb_cosine_bwd_rows", _p(t1), _p(t2), _p(kw), _p(stats), _p(dkw), R, D)


def vq_diagnostics(cos, stats, idx, hist, avg, ent):
    R, V = cos.shape
    _call("scb_vq_diagnostics", _p(cos), R, V, cos.stride(0), _p(stats), _p(idx), _p(hist), _p(avg), _p(ent))


def keyword_embed(emb, pos, idx, sot, eot, x0, keywords):
    B, K = idx.shape
    _call("scb_keyword_embed", _p(emb), _p(pos), _p(idx), sot, eot, B, K, emb.shape[1], _p(x0), _p(keywords))


def attention_small_bwd(qkv, dctx, dqkv, B, L, heads, hd, scale, causal):
    _call("scb_attention_small_bwd", _p(qkv), _DT[qkv.dtype], _p(dctx), _p(dqkv), B, L, heads, hd, scale, int(causal))


def act16_fwd(pre, act, out):
    _call("scb_act16_fwd", _p(pre), _DT[pre.dtype], act, _p(out), pre.numel())


def act_bwd(dy, pre, act, dx):
    _call("scb_act_bwd", _p(dy), _p(pre), _DT[pre.dtype], act, _p(dx), pre.numel())


def token_embed(emb, pos, tokens, x):
    B, L = tokens.shape
    _call("scb_token_embed", _p(emb), _p(pos), _p(tokens), B, L, emb.shape[1], emb.shape[0], _p(x))


def gather_rows(src, row, out):
    B, L, D = src.shape
    _call("scb_gather_rows", _p(src), _p(row), B, L, D, _p(out))


def softmax_rows(s, rows_per_batch, lens, cols, out, out_cols):
    """s fp32 [rows, ld]; lens int32 [rows / rows_per_batch] or None; out 16-bit [rows, out_ld]."""
    _call("scb_softmax_rows", _p(s), s.stride(0), s.shape[0], rows_per_batch, _p(lens), cols, _p(out), _DT[out.dtype], out.stride(0), out_cols)


# ------------------------------------------------------------------------------------------------ input side / pooling
def image_normalize(img_u8, mean, std, out=None):
    """uint8 [B, H, W, 3] (device) -> fp32 [B, 3, H, W], (x / 255 - mean) / std  (CLIP's ToTensor + Normalize)."""
    B, H, W, C = img_u8.shape
    assert C == 3 and img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    if out is None:
        out = torch.empty(B, 3, H, W, device=img_u8.device, dtype=torch.float32)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s_ = (ctypes.c_float * 3)(*[float(v) for v in std])
    _call("scb_image_normalize", _p(img_u8), B, H, W, m, s_, _p(out))
    return out


def pad_rows(packed, offsets, lens, tmax, out=None):
    """Ragged fp32 rows packed back to back (device) -> zero-padded [B, tmax]."""
    B = lens.numel()
    assert packed.dtype == torch.float32 and offsets.dtype == lens.dtype == torch.int64
    if out is None:
        out = torch.empty(B, tmax, device=packed.device, dtype=torch.float32)
    _call("scb_pad_rows", _p(packed), _p(offsets), _p(lens), B, tmax, _p(out))
    return out


def masked_mean_fwd(x, lens, out):
    B, T, D = x.shape
    _call("scb_masked_mean_fwd", _p(x), _p(lens), B, T, D, _p(out))
    return out


def masked_mean_bwd(dout, lens, dx):
    B, T, D = dx.shape
    _call("scb_masked_mean_bwd", _p(dout), _p(lens), B, T, D, _p(dx))
    return dx


def attentive_pool_fwd(align, mask, A, Bm, outA, outB):
    B, TA, TB = align.shape
    _call("scb_attentive_pool_fwd", _p(align), _p(mask), _p(A), _p(Bm), B, TA, TB, A.shape[1], Bm.shape[1], _p(outA), _p(outB))


def tanh_softmax_dim1(x, mask, y):
    B, TA, N = x.shape
    _call("scb_tanh_softmax_dim1", _p(x), _p(mask), B, TA, N, _p(y))
    return y


def relu_fwd(x, y):
    _call("scb_relu_fwd", _p(x), _p(y), x.numel())
    return y


def relu_bwd(dy, y, dx):
    _call("scb_relu_bwd", _p(dy), _p(y), _p(dx), y.numel())
    return dx
