"""Reference: avssl/module/pooling.py:8-390 (MeanPoolingLayer, AttentivePoolingLayer).  Neither is instantiated by the
reference's models (only exported from ``avssl/module/__init__.py:3``); they are part of the module surface SURVEY.md §8(b)
lists, so both run here as sm_100a kernels.

* ``MeanPoolingLayer``: projections on the TF32 tensor-core GEMM, the masked mean in ``scb_masked_mean_fwd/bwd`` (differentiable).
* ``AttentivePoolingLayer``: the two alignment GEMMs through ``scb_sgemm`` and everything after them (tanh, mask, row / column max,
  two softmaxes, two weighted sums) in ONE kernel per call (``scb_attentive_pool_fwd``); ``cal_batch_embedding`` uses
  ``scb_tanh_softmax_dim1``.  Forward only: no shipped configuration trains through it, so there is no backward kernel and
  a call that would need one raises.
"""
from typing import Tuple

import torch
from torch import nn

from speechclip_b200 import ops
from speechclip_b200.functional import LinearFn, MaskedMeanFn

__all__ = ["MeanPoolingLayer", "AttentivePoolingLayer"]


class MeanPoolingLayer(nn.Module):
    def __init__(self, in_dim: int = 0, out_dim: int = 0, bias: bool = True, pre_proj: bool = True, post_proj: bool = True):
        super().__init__()
        self.pre_proj = None
        self.post_proj = None
        if in_dim > 0 and out_dim > 0:
            if pre_proj:
                self.pre_proj = nn.Linear(in_dim, out_dim, bias=bias)
            if post_proj:
                self.post_proj = nn.Linear(in_dim if not pre_proj else out_dim, out_dim, bias=bias)

    def forward(self, x: torch.Tensor, x_len: torch.Tensor = None) -> torch.Tensor:
        """x [B, T, D], x_len [B] or None -> [B, D']."""
        if not x.is_cuda:
            raise RuntimeError("MeanPoolingLayer: CUDA tensors required (no CPU path)")
        if self.pre_proj is not None:
            x = LinearFn.apply(x, self.pre_proj.weight, self.pre_proj.bias)
        x = MaskedMeanFn.apply(x, x_len)
        if self.post_proj is not None:
            x = LinearFn.apply(x, self.post_proj.weight, self.post_proj.bias)
        return x


class AttentivePoolingLayer(nn.Module):
    def __init__(self, dim_A: int, dim_B: int, degraded: bool = False) -> None:
        super().__init__()
        self.dim_A, self.dim_B, self.degraded = dim_A, dim_B, degraded
        if not degraded:
            self.U = nn.Parameter(torch.randn(dim_A, dim_B))
        else:
            assert dim_A == dim_B
            self.U = nn.Parameter(torch.eye(dim_A), requires_grad=False)

    # ---------------------------------------------------------------------------------------------------------------
    def generate_input_msk(self, input_A_lens: torch.Tensor = None, input_B_lens: torch.Tensor = None, max_Alen: int = 1,
                           max_Blen: int = 1) -> torch.Tensor:
        """Additive mask [bsz, max_Alen, max_Blen] (0 = on, -inf = off), float64 like the reference's (pooling.py:88-146),
        without its python loop over the batch."""
        if input_A_lens is None and input_B_lens is None:
            raise ValueError("input_A_lens and input_B_lens cannot both be None")
        if input_A_lens is not None and input_B_lens is not None:
            assert input_A_lens.shape[0] == input_B_lens.shape[0], (
                "input_A_lens and input_B_lens must have same bsz, but got {} and {} instead".format(
                    input_A_lens.shape[0], input_B_lens.shape[0]))
        ref = input_A_lens if input_A_lens is not None else input_B_lens
        bsz, device = ref.shape[0], ref.device
        msk = torch.zeros((bsz, max_Alen, max_Blen), device=device, dtype=float)
        if input_A_lens is not None:
            la = input_A_lens.view(bsz).to(device)
            assert not (la == 0).any(), "Modality A has 0 length"
            msk.masked_fill_((torch.arange(max_Alen, device=device)[None, :] >= la[:, None])[:, :, None], float("-inf"))
        if input_B_lens is not None:
            lb = input_B_lens.view(bsz).to(device)
            assert not (lb == 0).any(), "Modality B has 0 length"
            msk.masked_fill_((torch.arange(max_Blen, device=device)[None, :] >= lb[:, None])[:, None, :], float("-inf"))
        return msk

    def _check(self, *tensors):
        for t in tensors:
            if t is not None and not t.is_cuda:
                raise RuntimeError("AttentivePoolingLayer: CUDA tensors required (no CPU path)")
        if torch.is_grad_enabled() and (self.U.requires_grad or any(t is not None and t.requires_grad for t in tensors)):
            raise NotImplementedError("AttentivePoolingLayer on B200 is forward only (no shipped configuration trains through it): "
                                      "call it under torch.no_grad()")

    @staticmethod
    def _expand_mask(msk, bsz, TA, TB, device):
        if msk is None:
            return None
        msk = msk.to(device=device, dtype=torch.float32)
        return msk.expand(bsz, TA, TB).contiguous()

    def _alignment(self, A: torch.Tensor, Bm: torch.Tensor) -> torch.Tensor:
        """A [n, dA, TA], Bm [n, dB, TB] fp32 -> A^T U B [n, TA, TB] (two strided fp32 GEMMs per pair)."""
        n, dA, TA = A.shape
        dB, TB = Bm.shape[1], Bm.shape[2]
        U = self.U.detach().float()
        au = torch.empty(TA, dB, device=A.device, dtype=torch.float32)
        align = torch.empty(n, TA, TB, device=A.device, dtype=torch.float32)
        for i in range(n):
            ops.sgemm(A[i].t(), U.t(), au)           # [TA, dA] x [dA, dB]
            ops.sgemm(au, Bm[i].t(), align[i])       # [TA, dB] x [dB, TB]
        return align

    @torch.no_grad()
    def _pooled(self, A, Bm, msk):
        A, Bm = A.detach().float().contiguous(), Bm.detach().float().contiguous()
        n, dA, TA = A.shape
        dB, TB = Bm.shape[1], Bm.shape[2]
        align = self._alignment(A, Bm)
        outA = torch.empty(n, dA, device=A.device, dtype=torch.float32)
        outB = torch.empty(n, dB, device=A.device, dtype=torch.float32)
        ops.attentive_pool_fwd(align, self._expand_mask(msk, n, TA, TB, A.device), A, Bm, outA, outB)
        return outA, outB

    def forward(self, input_A: torch.Tensor, input_B: torch.Tensor, intput_msk: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """input_A [bsz, dA, TA], input_B [bsz, dB, TB], mask [bsz, TA | 1, TB | 1] -> ([bsz, dA], [bsz, dB])  (pooling.py:335-390)."""
        assert len(input_A.shape) == 3, "input_A.shape must be (bsz_A,dim,seq_len)"
        assert len(input_B.shape) == 3, "input_B.shape must be (bsz_B,dim,seq_len)"
        assert input_A.shape[0] == input_B.shape[0], "input_A and input_B must have same bsz, but got {} and {} instead".format(
            input_A.shape[0], input_B.shape[0])
        if intput_msk is not None:
            assert input_A.shape[0] == intput_msk.shape[0], "input and intput_msk must have same bsz, but got {} and {} instead".format(
                input_A.shape[0], input_B.shape[0])
        self._check(input_A, input_B, intput_msk)
        outA, outB = self._pooled(input_A, input_B, intput_msk)
        return outA.squeeze(), outB.squeeze()   # (the reference squeezes: a batch of one loses its batch axis)

    def batch_forward(self, input_A: torch.Tensor, input_B: torch.Tensor, intput_msk: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every A against every B: ([bsz_A, bsz_B, dA], [bsz_A, bsz_B, dB])  (pooling.py:148-246)."""
        assert len(input_A.shape) == 3, "input_A.shape must be (bsz_A,dim,seq_len)"
        assert len(input_B.shape) == 3, "input_B.shape must be (bsz_B,dim,seq_len)"
        if intput_msk is not None:
            assert input_A.shape[0] == intput_msk.shape[0], "input and intput_msk must have same bsz, but got {} and {} instead".format(
                input_A.shape[0], intput_msk.shape[0])
        self._check(input_A, input_B, intput_msk)
        nA, nB = input_A.shape[0], input_B.shape[0]
        TA, TB = input_A.shape[2], input_B.shape[2]
        A = input_A.unsqueeze(1).expand(nA, nB, *input_A.shape[1:]).reshape(nA * nB, *input_A.shape[1:])
        Bm = input_B.unsqueeze(0).expand(nA, nB, *input_B.shape[1:]).reshape(nA * nB, *input_B.shape[1:])
        msk = None
        if intput_msk is not None:
            msk = intput_msk.to(torch.float32).expand(nA, TA, TB).unsqueeze(1).expand(nA, nB, TA, TB).reshape(nA * nB, TA, TB)
        outA, outB = self._pooled(A, Bm, msk)
        return outA.view(nA, nB, -1), outB.view(nA, nB, -1)

    @torch.no_grad()
    def cal_batch_embedding(self, input_A: torch.Tensor, input_B: torch.Tensor, intput_msk: torch.Tensor = None) -> torch.Tensor:
        """input_A [bsz, dA, TA], input_B [dB, N] (one vector per B instance), mask [bsz, TA, 1] -> [bsz, dA, N]  (pooling.py:248-333)."""
        assert len(input_A.shape) == 3, "input_A.shape must be (bsz,dim,seq_len)"
        assert len(input_B.shape) == 2, "input_B.shape must be (dim,total_data_pairs_count)"
        if intput_msk is not None:
            assert input_A.shape[0] == intput_msk.shape[0], "input and intput_msk must have same bsz, but got {} and {} instead".format(
                input_A.shape[0], input_B.shape[0])
            assert intput_msk.shape[2] == 1
        for t in (input_A, input_B, intput_msk):
            if t is not None and not t.is_cuda:
                raise RuntimeError("AttentivePoolingLayer: CUDA tensors required (no CPU path)")
        A = input_A.detach().float().contiguous()
        Bm = input_B.detach().float().contiguous()
        bsz, dA, TA = A.shape
        N = Bm.shape[1]
        ub = torch.empty(dA, N, device=A.device, dtype=torch.float32)
        ops.sgemm(self.U.detach().float(), Bm.t(), ub)                       # U B: [dA, dB] x [dB, N]
        align = torch.empty(bsz, TA, N, device=A.device, dtype=torch.float32)
        score = torch.empty_like(align)
        out = torch.empty(bsz, dA, N, device=A.device, dtype=torch.float32)
        msk = intput_msk.to(device=A.device, dtype=torch.float32).reshape(bsz, TA).contiguous() if intput_msk is not None else None
        for i in range(bsz):
            ops.sgemm(A[i].t(), ub.t(), align[i])                            # A^T (U B): [TA, dA] x [dA, N]
        ops.tanh_softmax_dim1(align, msk, score)
        for i in range(bsz):
            ops.sgemm(A[i], score[i].t(), out[i])                            # A score: [dA, TA] x [TA, N]
        return out
