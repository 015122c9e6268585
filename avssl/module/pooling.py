"""Reference: avssl/module/pooling.py:8-390 (MeanPoolingLayer, AttentivePoolingLayer).  Neither is instantiated by the
reference's models (grep: only exported from ``avssl/module/__init__.py:3``); they are part of the module surface SURVEY.md
§8(b) lists, so both run here as sm_100a kernels (forward; see each class)."""
from typing import Tuple

import torch
from torch import nn

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
        raise NotImplementedError


class AttentivePoolingLayer(nn.Module):
    def __init__(self, dim_A: int, dim_B: int, degraded: bool = False) -> None:
        super().__init__()
        self.dim_A, self.dim_B, self.degraded = dim_A, dim_B, degraded
        if not degraded:
            self.U = nn.Parameter(torch.randn(dim_A, dim_B))
        else:
            assert dim_A == dim_B
            self.U = nn.Parameter(torch.eye(dim_A), requires_grad=False)

    def forward(self, input_A: torch.Tensor, input_B: torch.Tensor, intput_msk: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError
