"""Reference: avssl/module/losses.py:126-245 (MaskedContrastiveLoss).  The whole loss — similarity matrix, same-id negative
mask, exp / row+column sums / log, and the gradients — runs in ``scb_infonce``.  The reference's MAX_EYE=256 buffers are
kept for state-dict parity only: the kernel builds the masks from ``index`` for any batch size."""
import math

import torch
from torch import nn

from speechclip_b200.functional import InfoNCEFn

MAX_EYE = 256


class MaskedContrastiveLoss(nn.Module):
    def __init__(self, temperature: float = 0.07, temperature_trainable: bool = False, margin: float = 0.0, dcl: bool = False,
                 a2b: bool = True, b2a: bool = True):
        super().__init__()
        assert a2b or b2a, "Cannot set both `a2b` and `b2a` to False."
        self.temperature_trainable = temperature_trainable
        self.margin = margin
        self.dcl = dcl
        self.a2b = a2b
        self.b2a = b2a
        if temperature_trainable:
            self.temperature = nn.Parameter(torch.ones([]) * math.log(1 / temperature))
        else:
            self.temperature = 1 / temperature
        eye_mat = torch.eye(MAX_EYE, dtype=torch.bool)
        self.register_buffer("eye_mat", eye_mat)
        self.register_buffer("neg_eye_mat", ~eye_mat)
        self.register_buffer("eye_mat_fl", eye_mat.type(torch.float))
        self._scb_arena_fn = None

    @property
    def current_temperature(self) -> float:
        if self.temperature_trainable:
            return float(self.temperature.data.detach().float().exp().item())
        return float(self.temperature)

    def forward(self, feat_A: torch.Tensor, feat_B: torch.Tensor, index: torch.LongTensor = None) -> torch.Tensor:
        assert feat_A.shape == feat_B.shape, (feat_A.shape, feat_B.shape)
        if index is not None:
            assert index.shape[0] == feat_A.shape[0], (index.shape, feat_A.shape)
            index = index.to(device=feat_A.device, dtype=torch.int64)
        arena = self._scb_arena_fn() if self._scb_arena_fn is not None else None
        if self.temperature_trainable:
            return InfoNCEFn.apply(feat_A, feat_B, index, self.temperature, 0.0, self.margin, self.dcl, self.a2b, self.b2a, arena)
        return InfoNCEFn.apply(feat_A, feat_B, index, None, self.temperature, self.margin, self.dcl, self.a2b, self.b2a, arena)
