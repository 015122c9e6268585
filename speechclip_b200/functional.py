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


def workspace(device, slot: int = 0) -> Workspace:
    """The device's scratch-buffer set.  ``slot`` > 0 names further, independent sets (runtime.TowerPipeline keeps the frozen
    towers of consecutive batches in alternating slots, so the tower of batch i + 1 never writes what batch i still reads)."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    key = idx if slot == 0 else (idx, slot)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = Workspace(torch.device("cuda", idx))
    return ws


class DropoutState:
    """Device-side RNG state of a module's train-mode dropout: int64 [seed, step] (scb_dropout_mask).  The seed follows
    ``torch.initial_seed()`` (Lightning's ``seed_everything``) and the process rank; ``advance`` bumps ``step`` with a kernel,
    so a captured CUDA graph draws fresh masks at every replay."""

    def __init__(self):
        self.state = None

    def get(self, device) -> torch.Tensor:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self.state is None or self.state.device != device:
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
            seed = (torch.initial_seed() * 0x9E3779B1 + rank * 0x85EBCA77 + 0x5bd1e995) % (2 ** 62)
            self.state = torch.tensor([seed, 0], dtype=torch.int64).to(device)
        return self.state

    def advance(self, device) -> torch.Tensor:
        st = self.get(device)
        ops.rng_advance(st)
        return st


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
        base = self.flat_g[0].data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.grad is not None and p.grad.data_ptr() == base + 4 * off:
                return 1
        return 0

    def backward_buffer(self) -> int:
        """Buffer index for the backward pass that is executing NOW: chosen when the first of our Functions runs in an autograd
        graph task and kept for the rest of that task (parameters adopt their views as .grad while the pass is still running,
        which must not flip the choice), so it never depends on how many forwards preceded this backward."""
        task = torch._C._current_graph_task_id()
        if task < 0 or task != getattr(self, "_task", None):
            self._task, self._task_buf = task, self.grad_buffer_index()
        return self._task_buf

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
    """avssl/module/weighted_sum.py:26-45.  hidden fp32 or fp16 [L, B*T, d] (frozen tower output, no grad)."""

    @staticmethod
    def forward(ctx, weights, hidden, B, T, normalize, arena, guard=None):
        """guard: the tower's graph-output tensor object that ``hidden`` views (its ``_scb_generation`` moves with every replay)."""
        _require_cuda(hidden, "WeightedSumLayer")
        d = hidden.shape[-1]
        out = torch.empty(B, T, d, device=hidden.device, dtype=torch.float32)
        ops.weighted_sum(hidden, weights, normalize, out32=out)
        ctx.save_for_backward(weights, hidden)
        ctx.meta = (B, T, normalize, arena)
        # under CUDA-graph replay ``hidden`` is the graph's output slab, overwritten by the tower's next replay (engine.GraphCache)
        ctx.hidden_obj, ctx.hidden_gen = guard, getattr(guard, "_scb_generation", None)
        return out

    @staticmethod
    def backward(ctx, dout):
        weights, hidden = ctx.saved_tensors
        B, T, normalize, arena = ctx.meta
        if getattr(ctx.hidden_obj, "_scb_generation", None) != ctx.hidden_gen:
            raise RuntimeError("WeightedSumLayer.backward: the speech tower ran again (CUDA-graph replay) before this backward and "
                               "overwrote the hidden states it needs; run backward before the next forward or set SCB_CUDA_GRAPHS=0")
        buf = arena.backward_buffer() if arena else 0
        d = hidden.shape[-1]
        if dout.stride(2) != 1 or dout.stride(1) != d:
            dout = dout.contiguous()
        gw = _grad_like(weights, arena, buf)
        gw.zero_()
        scratch = torch.empty(64, device=hidden.device, dtype=torch.float32)
        ops.weighted_sum_bwd(hidden, weights, normalize, dout, T, dout.stride(0), 0, scratch, gw, 1.0)
        return gw, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------- parallel branch
class ParallelBranchFn(torch.autograd.Function):
    """kwClip.py:1076-1108 on the [CLS] row (see speechclip_b200/head.py)."""

    @staticmethod
    def forward(ctx, audio_feat, kv_len, head: ParallelHead, arena, drop, *params):
        """drop: None (eval) or (p, rng_state int64[2]) — train-mode dropout of the encoder layer."""
        _require_cuda(audio_feat, "KW_ParallelBranch")
        p = dict(zip(PARAM_ORDER, params))
        audio_feat = audio_feat.contiguous()
        out, saved = head.cls_forward(workspace(audio_feat.device), p, audio_feat.detach(), kv_len, drop)
        ctx.head, ctx.arena, ctx.saved, ctx.params = head, arena, saved, params
        ctx.need_dfeat = audio_feat.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        head, arena, params = ctx.head, ctx.arena, ctx.params
        p = dict(zip(PARAM_ORDER, params))
        buf = arena.backward_buffer() if arena else 0
        g = {name: _grad_like(t, arena, buf) for name, t in p.items()}
        dfeat = head.cls_backward(workspace(dout.device), p, ctx.saved, dout.contiguous(), g, need_dfeat=ctx.need_dfeat)
        ctx.saved = None
        return (dfeat, None, None, None, None) + tuple(g[name] if p[name].requires_grad else None for name in PARAM_ORDER)


# ---------------------------------------------------------------------------------------------------- cascaded branch
class CascadedBranchFn(torch.autograd.Function):
    """kwClip.py:857-916 end to end (see speechclip_b200/cascaded.py).  ``rt`` carries the frozen pieces and switches:
    dict(vocab, text, bn_buffers, temp, sot, eot, training, need_grad[, drop = (p, rng_state)]).  Returns (text feature, keywords, cos, idx, stats); only the
    first output is differentiable."""

    @staticmethod
    def forward(ctx, audio_feat, kv_len, head, arena, rt, *params):
        from .cascaded import PARAM_ORDER as ORDER
        _require_cuda(audio_feat, "KW_CascadedBranch")
        p = dict(zip(ORDER, params))
        audio_feat = audio_feat.contiguous()
        need_grad = rt["need_grad"] and (audio_feat.requires_grad or any(t.requires_grad for t in params))
        feat, keywords, cos, idx, stats, saved = head.forward(
            workspace(audio_feat.device), p, audio_feat.detach(), kv_len, rt["bn_buffers"], rt["vocab"], rt["text"], rt["temp"],
            rt["sot"], rt["eot"], rt["training"], need_grad, rt.get("drop"))
        ctx.head, ctx.arena, ctx.saved, ctx.params, ctx.rt = head, arena, saved, params, rt
        ctx.need_dfeat = audio_feat.requires_grad
        ctx.mark_non_differentiable(keywords, cos, idx, stats)
        return feat, keywords, cos, idx, stats

    @staticmethod
    def backward(ctx, dfeat, *_):
        from .cascaded import PARAM_ORDER as ORDER
        head, arena, params, rt = ctx.head, ctx.arena, ctx.params, ctx.rt
        if ctx.saved.get("text") is None:
            raise RuntimeError("KW_CascadedBranch: backward through an eval-mode forward (no activations were kept)")
        p = dict(zip(ORDER, params))
        buf = arena.backward_buffer() if arena else 0
        g = {name: _grad_like(t, arena, buf) for name, t in p.items()}
        dx = head.backward(workspace(dfeat.device), p, ctx.saved, dfeat.contiguous(), g, rt["vocab"], rt["text"], need_dfeat=ctx.need_dfeat)
        ctx.saved = None
        return (dx, None, None, None, None) + tuple(g[name] if p[name].requires_grad else None for name in ORDER)


class KwBatchNormFn(torch.autograd.Function):
    """Kw_BatchNorm (eachKw, parallel) as a standalone op: kw_bn.py:96-125."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, training):
        _require_cuda(x, "Kw_BatchNorm")
        x = x.contiguous().float()
        B, K, W = x.shape
        y = torch.empty_like(x)
        mean = rstd = None
        if training:
            mean, rstd = torch.empty(K * W, device=x.device), torch.empty(K * W, device=x.device)
        ops.batchnorm_fwd(x, y, weight, bias, running_mean, running_var, mean, rstd, eps, momentum, training)
        ctx.save_for_backward(x, weight)
        ctx.stats = (mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        mean, rstd = ctx.stats
        if mean is None:
            raise RuntimeError("Kw_BatchNorm: backward through an eval-mode forward is not supported")
        dx, dg, db = torch.empty_like(x), torch.empty_like(weight), torch.empty_like(weight)
        ops.batchnorm_bwd(dy.contiguous().float(), x, weight, mean, rstd, dx, dg, db)
        return dx, dg, db, None, None, None, None, None


class VectorQuantizeFn(torch.autograd.Function):
    """SimpleVectorQuantizer on ready-made scores (my_vector_quantizer.py:66-135): returns (scores with masked ids = -inf,
    idx, stats); the caller builds the one-hot ``subword_prob``.  The gradient arriving at ``subword_prob`` is routed through
    softmax(scores / temp) as the hard straight-through estimator prescribes."""

    @staticmethod
    def forward(ctx, scores, mask_ids, temp):
        _require_cuda(scores, "SimpleVectorQuantizer")
        B, K, V = scores.shape
        Vp = (V + 3) // 4 * 4
        cos = torch.empty(B * K, Vp, device=scores.device, dtype=torch.float32)[:, :V]
        cos.copy_(scores.reshape(B * K, V))
        idx = torch.empty(B, K, device=scores.device, dtype=torch.int64)
        stats = torch.empty(B * K, 4, device=scores.device, dtype=torch.float32)
        ops.vq_forward(cos, None, None, mask_ids, temp, idx, stats)
        ctx.save_for_backward(cos, stats)
        ctx.temp = temp
        ctx.mark_non_differentiable(idx, stats)
        return cos, idx, stats

    @staticmethod
    def backward(ctx, dcos_unused, *_):
        raise RuntimeError("SimpleVectorQuantizer: differentiate through subword_prob (OneHotSTFn), not through the masked scores")


class OneHotSTFn(torch.autograd.Function):
    """subword_prob = one_hot(idx) + softmax(cos / temp) - softmax(cos / temp).detach()  (my_vector_quantizer.py:104-110)."""

    @staticmethod
    def forward(ctx, scores, cos, idx, stats, temp):
        B, K, V = scores.shape
        ctx.save_for_backward(cos, stats)
        ctx.temp, ctx.shape = temp, (B, K, V)
        return torch.zeros(B * K, V, device=scores.device).scatter_(1, idx.view(-1, 1), 1.0).view(B, K, V)

    @staticmethod
    def backward(ctx, dprob):
        cos, stats = ctx.saved_tensors
        B, K, V = ctx.shape
        g = torch.empty(B * K, cos.stride(0), device=cos.device, dtype=torch.float32)[:, :V]
        g.copy_(dprob.reshape(B * K, V))
        t2 = torch.empty(B * K, device=cos.device, dtype=torch.float32)
        ops.vq_backward(g, cos, stats, ctx.temp, t2)
        return g.reshape(B, K, V), None, None, None, None


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
        ctx.state = (a, b, ids, log_mult, float(fixed_mult), float(margin), dcl, a2b, b2a, scratch, arena)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        a, b, ids, log_mult, fixed_mult, margin, dcl, a2b, b2a, scratch, arena = ctx.state
        ctx.state = None
        buf = arena.backward_buffer() if arena else 0
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


# ---------------------------------------------------------------------------------------------------- small differentiable row ops
# (MeanPoolingLayer / MLPLayers of the reference's module surface: avssl/module/pooling.py:8-60, projections.py:6-29.  fp32 rows on
#  the TF32 tensor-core GEMM, the same helpers the trainable head uses.)
class LinearFn(torch.autograd.Function):
    """y = x W^T + b on fp32 rows [..., K] (nn.Linear)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from .head import linear
        _require_cuda(x, "Linear")
        x2 = x.contiguous().float().view(-1, x.shape[-1])
        y = torch.empty(x2.shape[0], weight.shape[0], device=x.device, dtype=torch.float32)
        linear(x2, weight, y, bias=bias)
        ctx.save_for_backward(x2, weight)
        ctx.shape, ctx.has_bias = x.shape, bias is not None
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        from .head import dgrad, wgrad
        x2, weight = ctx.saved_tensors
        ws = workspace(dy.device)
        dy2 = dy.contiguous().float().view(-1, weight.shape[0])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            dgrad(ws, "lin", dy2, weight, dx)
            dx = dx.view(ctx.shape)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(weight)
            wgrad(ws, "lin", dy2, x2, dw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(weight.shape[0], device=dy.device, dtype=torch.float32)
            ops.column_sum(dy2, db)
        return dx, dw, db


class MaskedMeanFn(torch.autograd.Function):
    """mean over the first len[b] frames of x [B, T, D] (pooling.py:52-56)."""

    @staticmethod
    def forward(ctx, x, lens):
        _require_cuda(x, "MeanPoolingLayer")
        x = x.contiguous().float()
        B, T, D = x.shape
        if lens is not None:
            lens = lens.to(device=x.device, dtype=torch.int64).contiguous()
        out = torch.empty(B, D, device=x.device, dtype=torch.float32)
        ops.masked_mean_fwd(x, lens, out)
        ctx.lens, ctx.shape = lens, (B, T, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        dx = torch.empty(ctx.shape, device=dout.device, dtype=torch.float32)
        ops.masked_mean_bwd(dout.contiguous().float(), ctx.lens, dx)
        return dx, None


class ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "ReLU")
        x = x.contiguous().float()
        y = ops.relu_fwd(x, torch.empty_like(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.relu_bwd(dy.contiguous().float(), y, torch.empty_like(y))


class DropoutFn(torch.autograd.Function):
    """Elementwise dropout with the counter-based masks of scb_dropout_mask; ``drop`` = (p, rng_state, site)."""

    @staticmethod
    def forward(ctx, x, drop):
        x = x.contiguous().float()
        ctx.drop = (drop[0], drop[1].clone(), drop[2])
        return ops.dropout_rows(x, torch.empty_like(x), ctx.drop)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous().float()
        return ops.dropout_rows(dy, torch.empty_like(dy), ctx.drop), None
