"""Fused optimizer step for the trainable head.

Reference: ``KWClipBase.configure_optimizers`` (kwClip.py:666-694) builds ``torch.optim.Adam(lr, weight_decay)`` and the
Trainer clips the global gradient norm to ``gradient_clip_val`` (spchclp_p.yaml:108) before every step.  Here both are ONE
pass over ONE flat buffer (``scb_adam_step``): the parameters are views into ``GradArena.flat_p`` and the backward kernels
wrote the gradients into ``GradArena.flat_g``.
"""
from __future__ import annotations

import torch

from . import ops
from .functional import GradArena


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected moments, eps outside the sqrt)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm: float = 0.0,
                 arena_fn=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        assert len(self.param_groups) == 1, "FusedAdam steps one flat buffer: a single param group"
        self.max_grad_norm = float(max_grad_norm)
        self.arena_fn = arena_fn  # the model's arena (shared with its backward kernels); None: build a private one
        self.arena = None
        self._m = self._v = self._sumsq = None
        self._step = 0

    def _ensure_arena(self):
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps[0].is_cuda:
            raise RuntimeError("FusedAdam: parameters must live on a CUDA device (no CPU path)")
        if self.arena_fn is not None:
            arena = self.arena_fn()
            assert {id(p) for p in arena.params} == {id(p) for p in ps}, "FusedAdam: optimizer params differ from the model arena"
        else:
            arena = self.arena
            if arena is None or not arena.intact():
                arena = GradArena(ps)
        if arena is not self.arena:
            if self.arena is None or self.arena.numel != arena.numel:
                self._m = None  # new layout: moments restart
            self.arena = arena
        if self._m is None:
            dev = self.arena.flat_p.device
            self._m = torch.zeros(self.arena.numel, device=dev, dtype=torch.float32)
            self._v = torch.zeros(self.arena.numel, device=dev, dtype=torch.float32)
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float64)

    def _flat_grad(self) -> torch.Tensor:
        """The flat gradient buffer the backward kernels filled; falls back to gathering .grad tensors that autograd
        cloned instead of adopting (correct, just one extra copy)."""
        a = self.arena
        g0 = a.params[0].grad
        buf = None
        for k in (0, 1):
            if g0 is not None and g0.data_ptr() == a.flat_g[k].data_ptr() + 4 * a.offsets[0]:
                buf = k
        if buf is None:
            buf = 0
        flat = a.flat_g[buf]
        base = flat.data_ptr()
        with torch.no_grad():
            for p, off in zip(a.params, a.offsets):
                view = flat[off:off + p.numel()]
                if p.grad is None:
                    view.zero_()
                elif p.grad.data_ptr() != base + 4 * off:
                    view.copy_(p.grad.reshape(-1))
        return flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_arena()
        g = self.param_groups[0]
        self._step += 1
        ops.adam_step(self.arena.flat_p, self._flat_grad(), self._m, self._v, self._sumsq, 1.0, self.max_grad_norm, float(g["lr"]),
                      g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step)
        return loss
