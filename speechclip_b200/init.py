"""Deterministic synthetic initialisation for the ``pretrained=False`` path.

There is no network in the build / bench environment, so the HuBERT and CLIP towers cannot be
downloaded (reference: ``speech_encoder_plus.py:385-395`` loads ``hubert_base_ls960.pt``,
``clip_official.py:50`` calls ``clip.load``).  Benchmarks and parity tests therefore run on
weights drawn here: every tensor is filled from its own ``torch.Generator`` seeded by
``seed`` and a CRC of the parameter NAME, so two modules with the same state-dict keys (the
CUDA-backed product modules and the CPU oracle) receive bit-identical weights regardless of
construction order.
"""
from __future__ import annotations

import math
import zlib

import torch
from torch import nn


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


@torch.no_grad()
def seeded_init_(module: nn.Module, seed: int = 7122) -> nn.Module:
    """Fill every floating-point parameter of ``module`` in place (fan-in scaled normals; non-trivial norms/biases)."""
    pending_g = []
    for name, p in module.named_parameters():
        if not p.is_floating_point() or p.dim() == 0:
            continue  # temperature / logit_scale keep their constructed value
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "weight_g":
            pending_g.append((name, p))
            continue
        if leaf == "cls":
            val = torch.randn(p.shape, generator=g)
        elif leaf == "weights":  # layer weighted-sum logits
            val = 0.5 * torch.randn(p.shape, generator=g)
        elif p.dim() == 1 and leaf == "weight":  # LayerNorm / GroupNorm scale
            val = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        elif p.dim() == 1:  # biases, class_embedding, mask_emb
            val = 0.05 * torch.randn(p.shape, generator=g)
        else:
            fan_in = p[0].numel()
            gain = math.sqrt(2.0) if "feature_extractor" in name else 1.0
            val = gain / math.sqrt(fan_in) * torch.randn(p.shape, generator=g)
        p.copy_(val.to(p.dtype))
    params = dict(module.named_parameters())
    for name, p in pending_g:  # weight_norm gain = ||v|| * (1 + 0.1 N): effective weight stays O(v)
        v = params[name[: -len("weight_g")] + "weight_v"]
        norm = v.float().pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
        p.copy_((norm * (1.0 + 0.1 * torch.randn(p.shape, generator=_gen(seed, name)))).to(p.dtype))
    return module
