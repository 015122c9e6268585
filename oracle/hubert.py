"""Oracle: fairseq HuBERT forward, as the reference drives it.  TEST INFRASTRUCTURE ONLY.

The arithmetic lives in a third-party dependency that is absent from
/root/reference: fairseq pinned at b5a039c292facba9c73f59ff34621ec131d82341
(``requirements.txt:6``).  This file restates its published algorithm
(``fairseq.models.hubert.hubert.HubertModel``,
``fairseq.models.wav2vec.wav2vec2.{ConvFeatureExtractionModel,TransformerEncoder,
TransformerSentenceEncoderLayer}``, ``fairseq.modules.MultiheadAttention``) and
anchors on the reference's own call sites:

* ``avssl/module/speech_encoder_plus.py:67-107``  customFunc_hubert_forward
* ``avssl/module/speech_encoder_plus.py:29-64``   patched TransformerEncoder.extract_features
* ``avssl/module/speech_encoder_plus.py:506-518`` preprocess_input
* ``avssl/module/speech_encoder_plus.py:520-634`` wrapper forward (lengths, weighted sum)

Parity unpinned against fairseq itself (not installable here); cross-checked in
``tests/test_oracle_crosscheck.py`` against ``transformers.HubertModel``.

State-dict key names equal fairseq's so a real ``hubert_base_ls960.pt`` /
SpeechCLIP ``.ckpt`` (prefix ``audio_encoder.encoder.``) loads unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

CONV_SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2  # fairseq conv_feature_layers


@dataclass
class HubertCfg:
    embed_dim: int = 768
    layers: int = 12
    heads: int = 12
    ffn_dim: int = 3072
    extractor_mode: str = "default"   # "default" = GroupNorm on conv0 only; "layer_norm" = LN after every conv
    layer_norm_first: bool = False    # False = post-LN (base); True = pre-LN (large)
    normalize_wav: bool = False       # fairseq task.cfg.normalize (True for large)
    conv_bias: bool = False
    pos_kernel: int = 128
    pos_groups: int = 16
    final_dim: int = 256              # unused head kept for checkpoint key parity

    @staticmethod
    def named(name: str) -> "HubertCfg":
        if name in ("hubert", "hubert_base"):
            return HubertCfg()
        if name == "hubert_large_ll60k":
            return HubertCfg(1024, 24, 16, 4096, "layer_norm", True, True, False, final_dim=768)
        if name == "tiny":  # test-sized, same structure as base
            return HubertCfg(64, 2, 4, 128, "default", False, False, False, 16, 4, 16)
        if name == "tiny_large":  # test-sized, same structure as large
            return HubertCfg(64, 2, 4, 128, "layer_norm", True, True, False, 16, 4, 16)
        raise KeyError(name)


class _TransposeLast(nn.Module):
    def forward(self, x):
        return x.transpose(-2, -1)


class ConvFeatureExtractor(nn.Module):
    """fairseq ConvFeatureExtractionModel: 7 strided Conv1d, total stride 320, receptive field 400."""

    def __init__(self, cfg: HubertCfg, conv_spec=CONV_SPEC):
        super().__init__()
        self.conv_layers = nn.ModuleList()
        in_d = 1
        for i, (dim, k, s) in enumerate(conv_spec):
            conv = nn.Conv1d(in_d, dim, k, stride=s, bias=cfg.conv_bias)
            if cfg.extractor_mode == "layer_norm":
                block = nn.Sequential(conv, nn.Dropout(0.0),
                                      nn.Sequential(_TransposeLast(), nn.LayerNorm(dim), _TransposeLast()),
                                      nn.GELU())
            elif i == 0:
                block = nn.Sequential(conv, nn.Dropout(0.0), nn.GroupNorm(dim, dim, affine=True), nn.GELU())
            else:
                block = nn.Sequential(conv, nn.Dropout(0.0), nn.GELU())
            self.conv_layers.append(block)
            in_d = dim

    def forward(self, x: torch.Tensor, collect: Optional[list] = None) -> torch.Tensor:
        x = x.unsqueeze(1)  # [B,1,Tw]
        for blk in self.conv_layers:
            x = blk(x)
            if collect is not None:
                collect.append(x)
        return x  # [B,512,T]


class _PosConv(nn.Module):
    """Conv1d(d,d,k=128,pad=64,groups=16) with weight_norm(dim=2): keys weight_g [1,1,k], weight_v [d,d/g,k], bias."""

    def __init__(self, d: int, k: int, groups: int):
        super().__init__()
        self.k, self.groups = k, groups
        self.weight_g = nn.Parameter(torch.ones(1, 1, k))
        self.weight_v = nn.Parameter(torch.randn(d, d // groups, k) * math.sqrt(4.0 / (k * d)))
        self.bias = nn.Parameter(torch.zeros(d))

    def effective_weight(self) -> torch.Tensor:
        v = self.weight_v
        norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()  # norm over every dim except 2
        return self.weight_g * v / norm

    def forward(self, x_bct: torch.Tensor) -> torch.Tensor:
        return F.conv1d(x_bct, self.effective_weight(), self.bias, padding=self.k // 2, groups=self.groups)


class _SelfAttention(nn.Module):
    """fairseq MultiheadAttention (self-attention, separate q/k/v/out Linear with bias)."""

    def __init__(self, d: int, heads: int):
        super().__init__()
        self.heads, self.hd = heads, d // heads
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)

    def forward(self, x: torch.Tensor, key_pad: Optional[torch.Tensor]) -> torch.Tensor:
        B, T, D = x.shape
        q = self.q_proj(x) * (self.hd ** -0.5)            # scaled AFTER projection
        k, v = self.k_proj(x), self.v_proj(x)
        q = q.view(B, T, self.heads, self.hd).transpose(1, 2)
        k = k.view(B, T, self.heads, self.hd).transpose(1, 2)
        v = v.view(B, T, self.heads, self.hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)                       # [B,H,T,T]
        if key_pad is not None and key_pad.any():
            s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
        p = torch.softmax(s.float(), dim=-1)
        o = (p @ v).transpose(1, 2).reshape(B, T, D)
        return self.out_proj(o)


class EncoderLayer(nn.Module):
    """fairseq TransformerSentenceEncoderLayer."""

    def __init__(self, cfg: HubertCfg):
        super().__init__()
        d = cfg.embed_dim
        self.layer_norm_first = cfg.layer_norm_first
        self.self_attn = _SelfAttention(d, cfg.heads)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, cfg.ffn_dim)
        self.fc2 = nn.Linear(cfg.ffn_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)

    def forward(self, x, key_pad):
        if self.layer_norm_first:
            x = x + self.self_attn(self.self_attn_layer_norm(x), key_pad)
            x = x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))
        else:
            x = self.self_attn_layer_norm(x + self.self_attn(x, key_pad))
            x = self.final_layer_norm(x + self.fc2(F.gelu(self.fc1(x))))
        return x


class TransformerEncoder(nn.Module):
    def __init__(self, cfg: HubertCfg):
        super().__init__()
        self.layer_norm_first = cfg.layer_norm_first
        self.pos_conv = nn.ModuleList([_PosConv(cfg.embed_dim, cfg.pos_kernel, cfg.pos_groups)])  # key: pos_conv.0.*
        self.pos_kernel = cfg.pos_kernel
        self.layers = nn.ModuleList([EncoderLayer(cfg) for _ in range(cfg.layers)])
        self.layer_norm = nn.LayerNorm(cfg.embed_dim)

    def extract_features(self, x: torch.Tensor, frame_pad: Optional[torch.Tensor]):
        """speech_encoder_plus.py:29-64 (eval mode: dropout / layerdrop inactive)."""
        if frame_pad is not None:
            x = x.masked_fill(frame_pad[:, :, None], 0.0)                      # :32-33 index_put(x, mask, 0)
        pc = self.pos_conv[0](x.transpose(1, 2))
        if self.pos_kernel % 2 == 0:
            pc = pc[:, :, :-1]                                                 # fairseq SamePad
        x = x + F.gelu(pc).transpose(1, 2)                                     # :35-37
        if not self.layer_norm_first:
            x = self.layer_norm(x)                                             # :39-40
        layer_results = [x]                                                    # :47
        for layer in self.layers:
            x = layer(x, frame_pad)                                            # :49-53
            layer_results.append(x)
        return x, layer_results


class HubertModel(nn.Module):
    def __init__(self, cfg: HubertCfg):
        super().__init__()
        self.cfg = cfg
        self.feature_extractor = ConvFeatureExtractor(cfg)
        self.layer_norm = nn.LayerNorm(512)
        self.post_extract_proj = nn.Linear(512, cfg.embed_dim)
        self.encoder = TransformerEncoder(cfg)
        # unused-but-present checkpoint tensors
        self.mask_emb = nn.Parameter(torch.zeros(cfg.embed_dim))
        self.final_proj = nn.Linear(cfg.embed_dim, cfg.final_dim)
        self.label_embs_concat = nn.Parameter(torch.zeros(504, cfg.final_dim))

    @staticmethod
    def frame_padding_mask(n_frames: int, sample_pad: torch.Tensor) -> torch.Tensor:
        """fairseq HubertModel.forward_padding_mask: a frame is padding iff ALL its samples are."""
        extra = sample_pad.size(1) % n_frames
        if extra > 0:
            sample_pad = sample_pad[:, :-extra]
        return sample_pad.view(sample_pad.size(0), n_frames, -1).all(-1)

    def custom_forward(self, source: torch.Tensor, sample_pad: Optional[torch.Tensor], collect: Optional[dict] = None):
        """speech_encoder_plus.py:67-107 with mask=None."""
        conv_outs = [] if collect is not None else None
        feats = self.feature_extractor(source, conv_outs).transpose(1, 2)      # :75-77  [B,T,512]
        feats = self.layer_norm(feats)                                         # :78
        frame_pad = None
        if sample_pad is not None:
            frame_pad = self.frame_padding_mask(feats.size(1), sample_pad)     # :81-82
        x = self.post_extract_proj(feats)                                      # :84-85
        out, layer_results = self.encoder.extract_features(x, frame_pad)       # :101
        if collect is not None:
            collect.update(conv_outs=conv_outs, features_ln=feats, post_proj=x, frame_pad=frame_pad)
        return {"x": out, "layer_results": layer_results, "frame_pad": frame_pad}


def preprocess_input(wavs: Sequence[torch.Tensor], normalize: bool):
    """speech_encoder_plus.py:506-518."""
    if normalize:
        wavs = [F.layer_norm(w, w.shape) for w in wavs]
    lens = torch.tensor([len(w) for w in wavs], dtype=torch.long)
    pad_mask = ~(torch.arange(int(lens.max())).unsqueeze(0) < lens.unsqueeze(1))
    padded = nn.utils.rnn.pad_sequence(list(wavs), batch_first=True)
    return padded, pad_mask


def feat_lengths(wav_lens: Sequence[int], n_frames: int, rate: int = 320) -> torch.Tensor:
    """speech_encoder_plus.py:602-611: clamp_max(round(len/320), T) with Python banker's round."""
    return torch.clamp_max(torch.tensor([round(l / rate) for l in wav_lens], dtype=torch.long), n_frames)


def conv_out_length(n: int) -> int:
    for _, k, s in CONV_SPEC:
        n = (n - k) // s + 1
    return n
