"""Reference: avssl/module/weighted_sum.py:8-45 — softmax-weighted sum of the hidden states (optional parameter-free
LayerNorm of each state first).  Forward/backward run in ``scb_weighted_sum_fwd/bwd``."""
from typing import List, Sequence, Union

import torch
from torch import nn

from speechclip_b200.functional import WeightedSumFn


class WeightedSumLayer(nn.Module):
    def __init__(self, n_weights: int, normalize_features: bool = False):
        super().__init__()
        self.n_weights = n_weights
        self.weights = nn.Parameter(torch.zeros((n_weights,), dtype=torch.float))
        self.normalize_features = normalize_features
        self._scb_arena_fn = None

    def forward(self, x: Union[Sequence[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        """x: list of n_weights tensors [B, T, d], or the stacked slab [n_weights, B, T, d] the HuBERT plan produces."""
        if isinstance(x, torch.Tensor):
            slab = x
        else:
            assert len(x) == self.n_weights, len(x)
            slab = torch.stack(list(x), dim=0)
        assert slab.shape[0] == self.n_weights, slab.shape
        L, B, T, d = slab.shape
        arena = self._scb_arena_fn() if self._scb_arena_fn is not None else None
        guard = getattr(x, "_scb_graph_output", None)  # set by the HuBERT wrapper when the slab is a CUDA-graph output buffer
        slab = slab.detach()
        if slab.dtype != torch.float16:  # fp16 slabs (post-LN tower) are read as they are; anything else as fp32
            slab = slab.float()
        return WeightedSumFn.apply(self.weights, slab.contiguous().view(L, B * T, d), B, T, self.normalize_features, arena, guard)
