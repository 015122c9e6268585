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


class SupConLoss(nn.Module):
    """Reference: avssl/module/losses.py:8-123 (supervised-contrastive / SimCLR loss).  Every shipped config selects
    ``MaskedContrastiveLoss`` (``cl_loss.type``, spchclp_*.yaml:71; ``SupConLoss`` appears there only as a comment), and the
    reference's own ``compute_loss`` calls the criterion as ``criterion(feat_A=, feat_B=, index=)`` (kwClip.py:1274-1290), which
    this class's ``forward(features, labels, mask)`` cannot accept — it is unreachable from the hot path.  The name and the
    constructor are kept so ``avssl.module`` imports and ``getattr(losses, ...)`` resolve; evaluating it raises."""

    def __init__(self, temperature=0.07, contrast_mode="all", base_temperature=0.07, learnable_temperature=True):
        super().__init__()
        self.learnable_temperature = learnable_temperature
        if learnable_temperature:
            self.temperature = nn.Parameter(torch.tensor([float(temperature)]))
        else:
            self.temperature = temperature
        self.contrast_mode = contrast_mode
        self.base_temperature = base_temperature

    @property
    def current_temperature(self):
        return self.temperature.item() if self.learnable_temperature else self.temperature

    def forward(self, features, labels=None, mask=None):
        raise NotImplementedError("SupConLoss has no sm_100a kernel: it is outside the SpeechCLIP hot path (no shipped config or "
                                  "call site of the reference can reach it); use MaskedContrastiveLoss")
