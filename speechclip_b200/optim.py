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
        self._pending_state = None  # torch-Adam layout loaded before the flat buffers exist (load_state_dict)

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
        if self._pending_state is not None:
            self._adopt_pending()

    # ------------------------------------------------------------------------------------------------- checkpoint state
    def _slices(self):
        """(index in param_groups[0]['params'], parameter, offset in the flat buffers) for every stepped parameter."""
        slot = {id(p): off for p, off in zip(self.arena.params, self.arena.offsets)}
        return [(i, p, slot[id(p)]) for i, p in enumerate(self.param_groups[0]["params"]) if id(p) in slot]

    def state_dict(self):
        """``torch.optim.Adam``'s layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``, same parameter order as the
        reference's ``configure_optimizers``, kwClip.py:666-694), so Lightning's ``optimizer_states`` round-trips and a reference
        checkpoint's Adam state loads.  The moments are copies of slices of the flat buffers."""
        sd = super().state_dict()
        if self._pending_state is not None and self._m is None:
            sd["state"] = {i: dict(st) for i, st in self._pending_state.items()}
        elif self._m is not None and self._step > 0:
            sd["state"] = {i: {"step": torch.tensor(float(self._step)),
                               "exp_avg": self._m[off:off + p.numel()].view(p.shape).clone(),
                               "exp_avg_sq": self._v[off:off + p.numel()].view(p.shape).clone()} for i, p, off in self._slices()}
        return sd

    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        self._pending_state = {int(i): st for i, st in state.items()} or None
        if self._pending_state is not None and self._m is not None:
            self._adopt_pending()

    def _adopt_pending(self):
        pend, self._pending_state = self._pending_state, None
        steps = []
        self._m.zero_()
        self._v.zero_()
        for i, p, off in self._slices():
            st = pend.get(i)
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"FusedAdam.load_state_dict: state {i} has shape {tuple(st['exp_avg'].shape)}, parameter {tuple(p.shape)}")
            self._m[off:off + p.numel()].copy_(st["exp_avg"].reshape(-1))
            self._v[off:off + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(st["step"]))
        self._step = max(steps) if steps else 0

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
        skipped = []
        with torch.no_grad():
            for p, off in zip(a.params, a.offsets):
                view = flat[off:off + p.numel()]
                if p.grad is None:
                    view.zero_()
                    skipped.append((off, p.numel()))
                elif p.grad.data_ptr() != base + 4 * off:
                    view.copy_(p.grad.reshape(-1))
        return flat, skipped

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_arena()
        g = self.param_groups[0]
        self._step += 1
        flat, skipped = self._flat_grad()
        # torch.optim.Adam leaves a parameter without gradient untouched (no decay, no moment update): keep its slices aside
        keep = [(off, n, [t[off:off + n].clone() for t in (self.arena.flat_p, self._m, self._v)]) for off, n in skipped]
        ops.adam_step(self.arena.flat_p, flat, self._m, self._v, self._sumsq, 1.0, self.max_grad_norm, float(g["lr"]),
                      g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step)
        for off, n, saved in keep:
            for t, v in zip((self.arena.flat_p, self._m, self._v), saved):
                t[off:off + n].copy_(v)
        return loss
