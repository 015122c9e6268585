"""Thin tensor-level wrappers over the C ABI: take torch CUDA tensors, pass raw pointers + sizes.

torch is used for device memory and the current stream only; every arithmetic op below runs in
libspeechclip_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import lib as _l

_DT = {torch.float32: _l.F32, torch.float16: _l.F16, torch.bfloat16: _l.BF16}


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _need(t: torch.Tensor, *dtypes):
    assert t.is_cuda, "CUDA tensor required (there is no CPU path)"
    assert t.dtype in dtypes, (t.dtype, dtypes)
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, act: int = _l.ACT_NONE,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.float16, out2: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """out[M,N] = act(alpha * a[M,K] @ w[N,K]^T + bias) + residual   (a, w 16-bit; fp32 accumulate on tcgen05)."""
    _need(a, torch.float16, torch.bfloat16)
    assert w.dtype == a.dtype and a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    g = _l.GemmArgs()
    g.a, g.a_inner, g.a_rows, g.a_row_stride, g.a_batch_stride = a.data_ptr(), K, M, a.stride(0), 0
    g.batch, g.m_per_batch = 1, M
    g.kb_per_tap, g.tap_row_shift, g.a_col0, g.a_group_cols = (K + 63) // 64, 0, 0, 0
    g.b, g.b_row_stride, g.b_group_stride, g.n, g.k, g.groups = w.data_ptr(), w.stride(0), 0, N, K, 1
    g.out, g.out_dtype, g.out_group_cols, g.ldc, g.out_batch_stride = out.data_ptr(), _DT[out.dtype], 0, out.stride(0), 0
    if out2 is not None:
        assert out2.shape == out.shape and out2.stride(0) == out.stride(0)
        g.out2, g.out2_dtype = out2.data_ptr(), _DT[out2.dtype]
    g.ab_format = _DT[a.dtype]
    if bias is not None:
        _need(bias, torch.float32)
        g.bias = bias.data_ptr()
    if residual is not None:
        assert residual.shape == out.shape and residual.stride(0) == out.stride(0)
        g.residual, g.residual_dtype = residual.data_ptr(), _DT[residual.dtype]
    g.act, g.alpha = act, alpha
    _l.check(_l.load().scb_gemm(ctypes.byref(g), _stream()), "scb_gemm")
    return out
