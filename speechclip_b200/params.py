"""Parameter containers for the frozen towers, keyed exactly like the upstream checkpoints.

The towers never run through torch: these modules only HOLD fp32 parameters under fairseq's / openai CLIP's
state-dict names so that ``state_dict()`` / ``load_state_dict()`` of the reference's checkpoints
work unchanged; ``speechclip_b200.engine`` compiles them into GEMM-layout fp16 plans.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
from torch import nn


class _Restore:
    """Set while a model is being rebuilt from a Lightning checkpoint (``load_from_checkpoint``): the checkpoint's
    ``state_dict`` carries every tower tensor, so the towers must not look for the separate fairseq / openai files that the
    pickled config's ``pretrained: true`` asks for (a released SpeechCLIP ``.ckpt`` loads on a machine that never had them)."""
    active = False


class restoring_from_checkpoint:
    def __enter__(self):
        self._prev, _Restore.active = _Restore.active, True

    def __exit__(self, *exc):
        _Restore.active = self._prev


def restoring() -> bool:
    return _Restore.active


class ParamTree(nn.Module):
    """A module tree built from dotted parameter names (numeric components become children named "0", "1", ...)."""

    @staticmethod
    def from_shapes(shapes: Dict[str, Tuple[int, ...]]) -> "ParamTree":
        root = ParamTree()
        for key, shape in shapes.items():
            parts = key.split(".")
            m = root
            for part in parts[:-1]:
                if part not in m._modules:
                    m.add_module(part, ParamTree())
                m = m._modules[part]
            m.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
        return root

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ParamTree only holds parameters; the forward runs in speechclip_b200.engine")


@dataclass
class HubertArch:
    embed_dim: int = 768
    layers: int = 12
    heads: int = 12
    ffn_dim: int = 3072
    extractor_layer_norm: bool = False   # fairseq extractor_mode == "layer_norm" (large)
    layer_norm_first: bool = False       # pre-LN encoder (large)
    normalize_wav: bool = False          # fairseq task.cfg.normalize (large)
    conv_bias: bool = False
    pos_kernel: int = 128
    pos_groups: int = 16
    final_dim: int = 256

    @staticmethod
    def named(name: str) -> "HubertArch":
        if name in ("hubert", "hubert_base"):
            return HubertArch()
        if name == "hubert_large_ll60k":
            return HubertArch(1024, 24, 16, 4096, True, True, True, False, final_dim=768)
        if name == "tiny":        # test-sized, base structure
            return HubertArch(64, 2, 4, 128, False, False, False, False, 16, 4, 16)
        if name == "tiny_large":  # test-sized, large structure
            return HubertArch(64, 2, 4, 128, True, True, True, False, 16, 4, 16)
        raise KeyError(name)


def hubert_param_shapes(a: HubertArch) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    spec = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
    cin = 1
    for i, (co, k, _) in enumerate(spec):
        s[f"feature_extractor.conv_layers.{i}.0.weight"] = (co, cin, k)
        if a.conv_bias:
            s[f"feature_extractor.conv_layers.{i}.0.bias"] = (co,)
        if a.extractor_layer_norm:
            s[f"feature_extractor.conv_layers.{i}.2.1.weight"] = (co,)
            s[f"feature_extractor.conv_layers.{i}.2.1.bias"] = (co,)
        elif i == 0:
            s["feature_extractor.conv_layers.0.2.weight"] = (co,)
            s["feature_extractor.conv_layers.0.2.bias"] = (co,)
        cin = co
    d = a.embed_dim
    s["layer_norm.weight"], s["layer_norm.bias"] = (512,), (512,)
    s["post_extract_proj.weight"], s["post_extract_proj.bias"] = (d, 512), (d,)
    s["encoder.pos_conv.0.weight_g"] = (1, 1, a.pos_kernel)
    s["encoder.pos_conv.0.weight_v"] = (d, d // a.pos_groups, a.pos_kernel)
    s["encoder.pos_conv.0.bias"] = (d,)
    for l in range(a.layers):
        p = f"encoder.layers.{l}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"], s[p + f"self_attn.{n}.bias"] = (d, d), (d,)
        s[p + "self_attn_layer_norm.weight"], s[p + "self_attn_layer_norm.bias"] = (d,), (d,)
        s[p + "fc1.weight"], s[p + "fc1.bias"] = (a.ffn_dim, d), (a.ffn_dim,)
        s[p + "fc2.weight"], s[p + "fc2.bias"] = (d, a.ffn_dim), (d,)
        s[p + "final_layer_norm.weight"], s[p + "final_layer_norm.bias"] = (d,), (d,)
    s["encoder.layer_norm.weight"], s["encoder.layer_norm.bias"] = (d,), (d,)
    # present in the upstream checkpoints, unused on this path
    s["mask_emb"] = (d,)
    s["final_proj.weight"], s["final_proj.bias"] = (a.final_dim, d), (a.final_dim,)
    s["label_embs_concat"] = (504, a.final_dim)
    return s


@dataclass
class ClipArch:
    image_size: int = 224
    patch: int = 32
    v_width: int = 768
    v_layers: int = 12
    v_heads: int = 12
    embed_dim: int = 512
    t_width: int = 512
    t_layers: int = 12
    t_heads: int = 8
    context: int = 77
    vocab: int = 49408

    @staticmethod
    def named(name: str) -> "ClipArch":
        if name == "ViT-B/32":
            return ClipArch()
        if name == "ViT-B/16":
            return ClipArch(patch=16)
        if name == "ViT-L/14":
            return ClipArch(224, 14, 1024, 24, 16, 768, 768, 12, 12)
        if name == "tiny":
            return ClipArch(32, 16, 64, 2, 4, 32, 32, 2, 4, 16, 64)
        if name == "tiny_c":      # tiny image tower + a text tower the cascaded branch can run (head_dim 16)
            return ClipArch(32, 16, 64, 2, 4, 32, 64, 2, 4, 16, 96)
        raise KeyError(name)


def clip_param_shapes(a: ClipArch) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def tower(prefix, width, layers):
        for l in range(layers):
            p = f"{prefix}resblocks.{l}."
            s[p + "attn.in_proj_weight"], s[p + "attn.in_proj_bias"] = (3 * width, width), (3 * width,)
            s[p + "attn.out_proj.weight"], s[p + "attn.out_proj.bias"] = (width, width), (width,)
            s[p + "ln_1.weight"], s[p + "ln_1.bias"] = (width,), (width,)
            s[p + "mlp.c_fc.weight"], s[p + "mlp.c_fc.bias"] = (4 * width, width), (4 * width,)
            s[p + "mlp.c_proj.weight"], s[p + "mlp.c_proj.bias"] = (width, 4 * width), (width,)
            s[p + "ln_2.weight"], s[p + "ln_2.bias"] = (width,), (width,)

    w, g = a.v_width, a.image_size // a.patch
    s["visual.conv1.weight"] = (w, 3, a.patch, a.patch)
    s["visual.class_embedding"] = (w,)
    s["visual.positional_embedding"] = (g * g + 1, w)
    s["visual.ln_pre.weight"], s["visual.ln_pre.bias"] = (w,), (w,)
    tower("visual.transformer.", w, a.v_layers)
    s["visual.ln_post.weight"], s["visual.ln_post.bias"] = (w,), (w,)
    s["visual.proj"] = (w, a.embed_dim)
    tower("transformer.", a.t_width, a.t_layers)
    s["token_embedding.weight"] = (a.vocab, a.t_width)
    s["positional_embedding"] = (a.context, a.t_width)
    s["ln_final.weight"], s["ln_final.bias"] = (a.t_width,), (a.t_width,)
    s["text_projection"] = (a.t_width, a.embed_dim)
    s["logit_scale"] = ()
    return s
