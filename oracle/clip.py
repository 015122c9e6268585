"""Oracle: openai CLIP towers as the reference drives them.  TEST INFRASTRUCTURE ONLY.

Third-party dependency absent from /root/reference: ``git+https://github.com/openai/CLIP.git``
(unpinned HEAD, ``requirements.txt:4``).  This restates its published ``clip/model.py``
(``VisionTransformer``, ``Transformer``, ``ResidualAttentionBlock``, ``QuickGELU``, fp32
``LayerNorm``) and anchors on the reference's call sites:

* ``avssl/module/clip_official.py:200-209``  encode_image
* ``avssl/module/clip_official.py:211-218``  encode_text
* ``avssl/module/clip_official.py:220-264``  encode_keywords

Parity unpinned against openai/CLIP itself (not installable here); cross-checked in
``tests/test_oracle_crosscheck.py`` against ``transformers.CLIPVisionModelWithProjection`` and
``CLIPTextModelWithProjection``.  State-dict keys equal openai's (prefix ``clip.model.`` in a ckpt).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn


@dataclass
class ClipCfg:
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
    def named(name: str) -> "ClipCfg":
        if name == "ViT-B/32":
            return ClipCfg()
        if name == "ViT-L/14":
            return ClipCfg(224, 14, 1024, 24, 16, 768, 768, 12, 12)
        if name == "tiny":
            return ClipCfg(32, 16, 64, 2, 4, 32, 32, 2, 4, 16, 64)
        if name == "tiny_c":
            return ClipCfg(32, 16, 64, 2, 4, 32, 64, 2, 4, 16, 96)
        raise KeyError(name)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d: int, heads: int, causal: bool, ctx: int):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)  # keys: attn.in_proj_weight/bias, attn.out_proj.*
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d, 4 * d))
        self.mlp.add_module("gelu", QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(4 * d, d))
        self.ln_2 = nn.LayerNorm(d)
        self.heads = heads
        self.causal = causal

    def attention(self, x):  # x [B,L,D]; written out rather than calling nn.MultiheadAttention.forward
        B, L, D = x.shape
        hd = D // self.heads
        qkv = x @ self.attn.in_proj_weight.t() + self.attn.in_proj_bias
        q, k, v = qkv.split(D, dim=-1)
        q = q.view(B, L, self.heads, hd).transpose(1, 2) * hd ** -0.5
        k = k.view(B, L, self.heads, hd).transpose(1, 2)
        v = v.view(B, L, self.heads, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if self.causal:
            s = s + torch.full((L, L), float("-inf"), dtype=s.dtype).triu_(1)
        o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, D)
        return self.attn.out_proj(o)

    def forward(self, x):
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, causal, ctx):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, causal, ctx) for _ in range(layers)])

    def forward(self, x, collect=None):
        for blk in self.resblocks:
            x = blk(x)
            if collect is not None:
                collect.append(x)
        return x


class VisionTransformer(nn.Module):
    def __init__(self, cfg: ClipCfg):
        super().__init__()
        w, g = cfg.v_width, cfg.image_size // cfg.patch
        scale = w ** -0.5
        self.conv1 = nn.Conv2d(3, w, cfg.patch, cfg.patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(w))
        self.positional_embedding = nn.Parameter(scale * torch.randn(g * g + 1, w))
        self.ln_pre = nn.LayerNorm(w)
        self.transformer = Transformer(w, cfg.v_layers, cfg.v_heads, False, g * g + 1)
        self.ln_post = nn.LayerNorm(w)
        self.proj = nn.Parameter(scale * torch.randn(w, cfg.embed_dim))

    def forward(self, x, collect=None):
        x = self.conv1(x)                                          # [B,W,G,G]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [B,G*G,W]
        cls = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], 1) + self.positional_embedding
        x = self.ln_pre(x)
        if collect is not None:
            collect.append(x)
        x = self.transformer(x, collect)
        x = self.ln_post(x[:, 0])
        return x @ self.proj


class CLIP(nn.Module):
    def __init__(self, cfg: ClipCfg):
        super().__init__()
        self.cfg = cfg
        self.visual = VisionTransformer(cfg)
        self.transformer = Transformer(cfg.t_width, cfg.t_layers, cfg.t_heads, True, cfg.context)
        self.token_embedding = nn.Embedding(cfg.vocab, cfg.t_width)
        self.positional_embedding = nn.Parameter(0.01 * torch.randn(cfg.context, cfg.t_width))
        self.ln_final = nn.LayerNorm(cfg.t_width)
        self.text_projection = nn.Parameter(cfg.t_width ** -0.5 * torch.randn(cfg.t_width, cfg.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        x = self.token_embedding(text) + self.positional_embedding
        x = self.ln_final(self.transformer(x))
        return x[torch.arange(x.shape[0]), text.argmax(-1)] @ self.text_projection

    def encode_keywords(self, keywords, keyword_num, sot_token, eot_token):
        """clip_official.py:220-264: inject K keyword vectors at positions 1..K, read position K+1."""
        B = keywords.size(0)
        text = torch.zeros(B, self.cfg.context, dtype=torch.long)
        text[:, 0] = sot_token
        text[:, keyword_num + 1] = eot_token
        x = self.token_embedding(text).clone()
        x[:, 1:1 + keyword_num] = keywords
        x = x + self.positional_embedding
        x = self.ln_final(self.transformer(x))
        return x[:, 1 + keyword_num] @ self.text_projection
