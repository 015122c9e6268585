"""torch.autograd glue: each Function's forward/backward is a kernel sequence over the C ABI.

These are what make the CUDA path a drop-in under ``LightningModule.training_step`` + ``loss.backward()``:
the reference relies on autograd through torch library ops; here autograd only carries tensors between
our own forward/backward kernel sequences.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .engine import Workspace
from .head import PARAM_ORDER, ParallelHead

_WS: Dict[int, Workspace] = {}


def workspace(device) -> Workspace:
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    ws = _WS.get(idx)
    if ws is None:
        ws = _WS[idx] = Workspace(torch.device("cuda", idx))
    return ws


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required — the sm_100a extension is the only implementation of this path "
                           "(there is no CPU / eager fallback)")


class GradArena:
    """Flat fp32 storage for the trainable head: parameters are views into ``flat_p`` and the backward kernels write
    gradients into views of ``flat_g`` at the same offsets, so clip + Adam is one pass over one buffer (optim.py)."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = list(params)
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4  # keep every slot 16-byte aligned
        self.numel = n
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = [torch.zeros(n, device=dev, dtype=torch.float32) for _ in range(2)]
        self._slot = {id(p): i for i, p in enumerate(self.params)}
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat_p[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view

    def intact(self) -> bool:
        base = self.flat_p.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(self.params, self.offsets))

    def grad_buffer_index(self) -> int:
        """The flat gradient buffer that does NOT alias the live .grad tensors (so autograd's ``grad += new`` is safe)."""
        g = self.params[0].grad
        if g is not None and g.data_ptr() == self.flat_g[0].data_ptr() + 4 * self.offsets[0]:
            return 1
        return 0

    def grad_view(self, p: torch.nn.Parameter, buf: int) -> torch.Tensor:
        off = self.offsets[self._slot[id(p)]]
        return self.flat_g[buf][off:off + p.numel()].view(p.shape)

    def has(self, p) -> bool:
        return id(p) in self._slot


def _grad_like(p: torch.Tensor, arena: Optional[GradArena], buf: int) -> torch.Tensor:
    if arena is not None and arena.has(p):
        return arena.grad_view(p, buf)
    return torch.empty_like(p)


# ---------------------------------------------------------------------------------------------------- weighted sum
class WeightedSumFn(torch.autograd.Function):
    """avssl/module/weighted_sum.py:26-45.  hidden fp32 [L, B*T, d] (frozen tower output, no grad)."""

    @staticmethod
    def forward(ctx, weights, hidden, B, T, normalize, arena):
        _require_cuda(hidden, "WeightedSumLayer")
        d = hidden.shape[-1]
        out = torch.empty(B, T, d, device=hidden.device, dtype=torch.float32)
        ops.weighted_sum(hidden, weights, normalize, out32=out)
        ctx.save_for_backward(weights, hidden)
        ctx.meta = (B, T, normalize, arena, arena.grad_buffer_index() if arena else 0)
        return out

    @staticmethod
    def backward(ctx, dout):
        weights, hidden = ctx.saved_tensors
        B, T, normalize, arena, buf = ctx.meta
        d = hidden.shape[-1]
        if dout.stride(2) != 1 or dout.stride(1) != d:
            dout = dout.contiguous()
        gw = _grad_like(weights, arena, buf)
        gw.zero_()
        scratch = torch.empty(64, device=hidden.device, dtype=torch.float32)
        ops.weighted_sum_bwd(hidden, weights, normalize, dout, T, dout.stride(0), 0, scratch, gw, 1.0)
        return gw, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------- parallel branch
class ParallelBranchFn(torch.autograd.Function):
    """kwClip.py:1076-1108 on the [CLS] row (see speechclip_b200/head.py)."""

    @staticmethod
    def forward(ctx, audio_feat, kv_len, head: ParallelHead, arena, *params):
        _require_cuda(audio_feat, "KW_ParallelBranch")
        p = dict(zip(PARAM_ORDER, params))
        audio_feat = audio_feat.contiguous()
        out, saved = head.cls_forward(workspace(audio_feat.device), p, audio_feat.detach(), kv_len)
        ctx.head, ctx.arena, ctx.saved, ctx.params = head, arena, saved, params
        ctx.need_dfeat = audio_feat.requires_grad
        ctx.buf = arena.grad_buffer_index() if arena else 0
        return out

    @staticmethod
    def backward(ctx, dout):
        head, arena, params = ctx.head, ctx.arena, ctx.params
        p = dict(zip(PARAM_ORDER, params))
        buf = ctx.buf
        g = {name: _grad_like(t, arena, buf) for name, t in p.items()}
        dfeat = head.cls_backward(workspace(dout.device), p, ctx.saved, dout.contiguous(), g, need_dfeat=ctx.need_dfeat)
        ctx.saved = None
        return (dfeat, None, None, None) + tuple(g[name] if p[name].requires_grad else None for name in PARAM_ORDER)


# ---------------------------------------------------------------------------------------------------- L2 normalise
class L2NormFn(torch.autograd.Function):
    """x / x.norm(dim=-1, keepdim=True)  (kwClip.py:1436,1451-1453)."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "l2_normalize")
        x = x.contiguous().float()
        y = torch.empty_like(x)
        norms = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        ops.l2norm(x, y, norms)
        ctx.save_for_backward(y, norms)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norms = ctx.saved_tensors
        dx = torch.empty_like(y)
        ops.l2norm_bwd(dy.contiguous(), y, norms, dx)
        return dx


# ---------------------------------------------------------------------------------------------------- masked InfoNCE
class InfoNCEFn(torch.autograd.Function):
    """avssl/module/losses.py:185-245.  Forward leaves the logits in a per-call scratch; backward turns them into gradients."""

    @staticmethod
    def forward(ctx, feat_a, feat_b, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a, arena):
        _require_cuda(feat_a, "MaskedContrastiveLoss")
        a, b = feat_a.contiguous().float(), feat_b.contiguous().float()
        B = a.shape[0]
        loss = torch.empty((), device=a.device, dtype=torch.float32)
        scratch = torch.empty(ops.infonce_scratch_bytes(B), device=a.device, dtype=torch.uint8)
        if ids is not None:
            ids = ids.contiguous()
        ops.infonce(a, b, ids, log_mult, float(fixed_mult), float(margin), dcl, a2b, b2a, scratch, phase=1, loss=loss)
        ctx.state = (a, b, ids, log_mult, float(fixed_mult), float(margin), dcl, a2b, b2a, scratch, arena,
                     arena.grad_buffer_index() if arena else 0)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        a, b, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a, scratch, arena, buf = ctx.state
        ctx.state = None
        if scratch is None:
            raise RuntimeError("MaskedContrastiveLoss: backward called twice (the logits scratch was consumed)")
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_t = log_mult is not None and ctx.needs_input_grad[3]
        dA = torch.empty_like(a) if need_a else None
        dB = torch.empty_like(b) if need_b else None
        dT = None
        if need_t:
            dT = _grad_like(log_mult, arena, buf)
            dT.zero_()
        ops.infonce(a, b, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a, scratch, phase=2, upstream_dev=dloss.contiguous().float(),
                    dA=dA, dB=dB, dlog_mult=dT)
        return dA, dB, None, dT, None, None, None, None, None, None
