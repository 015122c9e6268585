"""Keyword BatchNorm (reference: avssl/module/speechclip_c_modules/kw_bn.py:8-153).

The shipped cascaded configs use ``type: eachKw, parallel: true``: one ``BatchNorm1d`` over kw_dim * kw_num features in
(dim, keyword) order, initialised from the token-embedding statistics.  Parameters and buffers live under the reference's
state-dict keys (``bn_layer.weight / bias / running_mean / running_var / num_batches_tracked``); the arithmetic runs in
``scb_batchnorm_fwd / _bwd``.  The non-parallel and ``same`` variants are not used by any shipped config."""
import logging

import torch
from torch import nn

from speechclip_b200.functional import KwBatchNormFn

logger = logging.getLogger(__name__)

__all__ = ["Kw_BatchNorm"]


class _BatchNormParams(nn.Module):
    """Parameter / buffer holder with ``nn.BatchNorm1d``'s state-dict layout (never called)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Kw_BatchNorm(nn.Module):
    def __init__(self, kw_num: int, kw_dim: int, batchnorm_type: str, init_bias: torch.Tensor, init_scale: torch.Tensor,
                 std_scale: int = 1, learnable: bool = True, parallel: bool = False) -> None:
        super().__init__()
        self.batchnorm_type, self.kw_num, self.kw_dim = batchnorm_type, kw_num, kw_dim
        self.std_scale, self.learnable, self.parallel = std_scale, learnable, parallel
        if not (batchnorm_type == "eachKw" and parallel):
            raise NotImplementedError("Kw_BatchNorm on B200: only type=eachKw with parallel=true (every shipped cascaded config)")
        self.bn_layer = _BatchNormParams(kw_dim * kw_num)
        if not isinstance(self.std_scale, list):
            self.std_scale = [self.std_scale] * self.kw_num
        self.init_bn(init_bias, init_scale)
        logger.info("Initialize BatchNorm({}) weight and bias learnable=({}) with token embeddings w/ scale={}, parallel=({})".format(
            self.batchnorm_type, self.learnable, self.std_scale, self.parallel))

    def init_bn(self, init_bias: torch.Tensor, init_scale: torch.Tensor) -> None:
        # kw_bn.py:77-83: the per-dim vectors are tiled kw_num times over the (dim, kw)-ordered feature axis, as the reference does
        self.bn_layer.weight.data.copy_((init_scale * self.std_scale[0]).repeat(self.kw_num))
        self.bn_layer.bias.data.copy_(init_bias.repeat(self.kw_num))
        self.bn_layer.weight.requires_grad = self.learnable
        self.bn_layer.bias.requires_grad = self.learnable

    def forward(self, keywords: torch.Tensor, seq_lens: torch.Tensor = None) -> torch.Tensor:
        assert keywords.dim() == 3
        assert keywords.shape[2] == self.kw_dim
        if seq_lens is not None:
            raise NotImplementedError("Kw_BatchNorm(seq_lens=...) belongs to the `same` variant, which no shipped config uses")
        assert keywords.shape[1] == self.kw_num
        bn = self.bn_layer
        y = KwBatchNormFn.apply(keywords, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, self.training)
        if self.training:
            bn.num_batches_tracked += 1
        return y
