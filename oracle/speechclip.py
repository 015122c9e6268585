"""Oracle: the SpeechCLIP contrastive forward / loss around the two towers.  TEST INFRASTRUCTURE ONLY.

Restates, in plain fp32 torch on CPU, the reference's own (torch-only) code; every function here
is PINNED by fixtures that ``tests/golden/make_golden.py`` generated from the reference files
themselves (imported from /root/reference in the build container):

* weighted_sum            ``avssl/module/weighted_sum.py:26-45``
* keypadding_mask         ``avssl/util/data_utils.py:4-20``
* BranchEncoder           ``avssl/module/kw_modules/TransformerModels.py:48-96`` (nn.TransformerEncoder,
                          1 post-LN layer + final LN)
* ParallelBranch          ``avssl/model/kwClip.py:1004-1108``
* masked_contrastive_loss ``avssl/module/losses.py:185-245`` (MAX_EYE cap at :126 lifted: eye(B))
* mutual_retrieval        ``avssl/module/retrieval.py:6-121``
* linear_warmup_decay     ``avssl/optim/scheduler.py:22-38``
* SpeechClipOracle        ``avssl/model/kwClip.py:1111-1478`` + ``speech_encoder_plus.py:520-634``
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from . import clip as oclip
from . import hubert as ohubert


def weighted_sum(weights: torch.Tensor, hidden: Sequence[torch.Tensor], normalize: bool = False) -> torch.Tensor:
    w = torch.softmax(weights, dim=0)
    x = torch.stack(list(hidden), dim=0)
    if normalize:
        x = F.layer_norm(x, (x.shape[-1],))
    return (w.view(-1, 1, 1, 1) * x).sum(0)


def keypadding_mask(max_length: int, lens: torch.Tensor) -> torch.Tensor:
    """True = padding."""
    return torch.arange(max_length)[None, :] >= lens[:, None]


class BranchEncoder(nn.Module):
    """n post-LN (or pre-LN) torch TransformerEncoderLayers + final LayerNorm; keys ``model.layers.N.*``, ``model.norm.*``."""

    def __init__(self, n_layers=1, d_model=768, nhead=8, dim_feedforward=3072, dropout=0.1, activation="gelu",
                 layer_norm_eps=1e-5, batch_first=True, norm_first=False):
        super().__init__()
        assert activation == "gelu" and batch_first
        layer = nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, layer_norm_eps,
                                           batch_first, norm_first)
        self.model = nn.TransformerEncoder(layer, n_layers, nn.LayerNorm(d_model, eps=1e-5), enable_nested_tensor=False)
        self.nhead, self.norm_first = nhead, norm_first

    # ``masks`` (train mode, TransformerModels.py:55-75 -> nn.TransformerEncoderLayer(dropout=p)): multiplicative dropout masks
    # (0 or 1/(1-p)) for the four dropouts of the layer, given for query row 0 only — the row the branch consumes (kwClip.py:1103);
    # keys "attn" [B, H, L], "dropout1" [B, D], "ffn" [B, F], "dropout2" [B, D]; masks with one more dimension cover every row
    # ("attn" [B, H, L, L], ...).  torch draws them from its own generator inside library code, which cannot be replayed
    # elsewhere; the oracle takes the masks as inputs so both sides evaluate the same realisation.  Pinned by
    # tests/golden/ref_branch_train_dropout.npz (the reference's module in train mode with injected masks).
    @staticmethod
    def _row0(x, mask):
        if mask is None:
            return x
        if mask.dim() == x.dim():
            return x * mask
        m = torch.ones_like(x)
        m[:, 0] = mask
        return x * m

    def _attn(self, lyr, x, key_pad, masks=None):
        B, L, D = x.shape
        hd = D // self.nhead
        qkv = x @ lyr.self_attn.in_proj_weight.t() + lyr.self_attn.in_proj_bias
        q, k, v = qkv.split(D, -1)
        q = q.view(B, L, self.nhead, hd).transpose(1, 2) * hd ** -0.5
        k = k.view(B, L, self.nhead, hd).transpose(1, 2)
        v = v.view(B, L, self.nhead, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
        p = torch.softmax(s, -1)
        if masks is not None:
            if masks["attn"].dim() == 4:
                p = p * masks["attn"]
            else:
                pm = torch.ones_like(p)
                pm[:, :, 0, :] = masks["attn"]
                p = p * pm
        o = (p @ v).transpose(1, 2).reshape(B, L, D)
        return o @ lyr.self_attn.out_proj.weight.t() + lyr.self_attn.out_proj.bias

    def _layer(self, lyr, x, key_pad, masks=None):
        m = masks or {}
        ff = lambda y: self._row0(lyr.linear2(self._row0(F.gelu(lyr.linear1(y)), m.get("ffn"))), m.get("dropout2"))
        if self.norm_first:
            x = x + self._row0(self._attn(lyr, lyr.norm1(x), key_pad, masks), m.get("dropout1"))
            return x + ff(lyr.norm2(x))
        x = lyr.norm1(x + self._row0(self._attn(lyr, x, key_pad, masks), m.get("dropout1")))
        return lyr.norm2(x + ff(x))

    def forward(self, src, key_padding_mask, hidden: Optional[list] = None, masks=None):
        x = src
        for lyr in self.model.layers:
            if hidden is not None:
                hidden.append(x)
            x = self._layer(lyr, x, key_padding_mask, masks)
        if hidden is not None:
            hidden.append(x)
        return self.model.norm(x)


class ParallelBranch(nn.Module):
    def __init__(self, d_model: int, out_dim: int, **transformer_args):
        super().__init__()
        self.self_att = BranchEncoder(d_model=d_model, **transformer_args)
        self.cls = nn.Parameter(torch.randn(1, 1, d_model))
        self.linear_proj = nn.Linear(d_model, out_dim)

    def _src(self, audio_feat, audio_len):
        B, T = audio_feat.shape[:2]
        src = torch.cat([self.cls.expand(B, 1, -1), audio_feat], 1)          # kwClip.py:1093-1094
        return src, keypadding_mask(T + 1, audio_len + 1)                     # :1096-1099

    def forward(self, audio_feat, audio_len, masks=None):
        src, kpm = self._src(audio_feat, audio_len)
        out = self.self_att(src, kpm, masks=masks)
        return self.linear_proj(out[:, 0])                                    # :1103-1106

    def extract_hidden_states(self, audio_feat, audio_len):
        src, kpm = self._src(audio_feat, audio_len)
        hidden: list = []
        self.self_att(src, kpm, hidden)
        return tuple(h[:, 1:] for h in hidden)                                # kwClip.py:1071-1073


class AttentionAndNorm(nn.Module):
    """``MultiheadAttentionAndNorm`` (TransformerModels.py:99-135): LN(MHA(src, src, src) + src); keys
    ``multihead_attn_layer.*``, ``attentionBlock_Norm.*``."""

    def __init__(self, d_model=768, nhead=1, dropout=0.1, layer_norm_eps=1e-5, batch_first=True, **_):
        super().__init__()
        assert batch_first
        self.multihead_attn_layer = nn.MultiheadAttention(d_model, num_heads=nhead, dropout=dropout, batch_first=True)
        self.attentionBlock_Norm = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.nhead = nhead

    def forward(self, src, key_padding_mask, attn_mask_rows=None):
        """attn_mask_rows: train-mode attention-dropout mask (0 or 1/(1-p)) [B, H, K, L] for the first K query rows (the keyword
        rows the cascaded branch consumes, kwClip.py:878-882); None = eval."""
        m = self.multihead_attn_layer
        B, L, D = src.shape
        hd = D // self.nhead
        qkv = src @ m.in_proj_weight.t() + m.in_proj_bias
        q, k, v = qkv.split(D, -1)
        q = q.view(B, L, self.nhead, hd).transpose(1, 2) * hd ** -0.5
        k = k.view(B, L, self.nhead, hd).transpose(1, 2)
        v = v.view(B, L, self.nhead, hd).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)).masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
        p = torch.softmax(s, -1)
        if attn_mask_rows is not None:
            pm = torch.ones_like(p)
            pm[:, :, :attn_mask_rows.shape[2], :] = attn_mask_rows
            p = p * pm
        o = (p @ v).transpose(1, 2).reshape(B, L, D)
        return self.attentionBlock_Norm(o @ m.out_proj.weight.t() + m.out_proj.bias + src)


class KwBatchNorm(nn.Module):
    """``Kw_BatchNorm`` with batchnorm_type=eachKw, parallel=True (kw_bn.py:97-123): one BatchNorm1d over kw_dim * kw_num
    features ordered (dim, kw); key ``bn_layer.*``."""

    def __init__(self, kw_num, kw_dim, init_bias, init_scale, std_scale=1.0, learnable=True):
        super().__init__()
        self.kw_num, self.kw_dim = kw_num, kw_dim
        self.bn_layer = nn.BatchNorm1d(kw_dim * kw_num)
        with torch.no_grad():
            self.bn_layer.weight.copy_((init_scale * std_scale).repeat(kw_num))   # (sic) the reference tiles per-dim values
            self.bn_layer.bias.copy_(init_bias.repeat(kw_num))                    # over a (dim, kw)-ordered axis
        self.bn_layer.weight.requires_grad = learnable
        self.bn_layer.bias.requires_grad = learnable

    def forward(self, keywords):
        B = keywords.shape[0]
        x = keywords.permute(0, 2, 1).reshape(B, -1)
        x = self.bn_layer(x)
        return x.reshape(B, self.kw_dim, self.kw_num).permute(0, 2, 1)


def simple_vector_quantizer(x, temp, training, prob_msk=(0, 2, 3), force_idx=None):
    """``SimpleVectorQuantizer.forward`` (my_vector_quantizer.py:64-165) with use_gumbel=False, hard=True, time_first=True.
    ``force_idx`` (test hook, not in the reference): use these ids instead of the argmax, so that a run whose upstream
    features differ in the last bits can be compared downstream even where two scores are within rounding of each other."""
    bsz, tsz, fsz = x.shape
    x = x.reshape(bsz * tsz, fsz).clone()
    for i in prob_msk:
        x[:, i] = x[:, i] + float("-inf")
    k = x.argmax(-1) if force_idx is None else force_idx.reshape(-1)
    hard_x = torch.zeros_like(x).scatter_(-1, k.view(-1, 1), 1.0)
    hard_probs = hard_x.float().mean(0)
    result = {"num_vars": fsz}
    result["code_perplexity"] = torch.exp(-torch.sum(hard_probs * torch.log(hard_probs + 1e-7), dim=-1)).sum()
    avg_probs = torch.softmax(x.float(), dim=-1).mean(0)
    probs_per_t = torch.softmax(x.view(bsz, tsz, -1), dim=-1).permute(1, 0, 2)
    result["ent_per_t"] = (-torch.sum(probs_per_t * torch.log(probs_per_t + 1e-9), dim=-1)).mean(-1)
    result["prob_perplexity"] = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-7), dim=-1)).sum()
    result["temp"] = float(temp)
    if training:
        soft = torch.softmax(x / temp, dim=-1)
        y = hard_x + soft - soft.detach()
    else:
        y = hard_x
    result["subword_prob"] = y.view(bsz, tsz, -1)
    result["diversity_loss"] = (fsz - result["prob_perplexity"]) / fsz
    result["targets"] = y.argmax(-1).view(bsz, tsz, 1).detach()
    return result


class CascadedBranch(nn.Module):
    """``KW_CascadedBranch`` (kwClip.py:697-916) with MultiheadAttentionAndNorm, eachKw/parallel BatchNorm, SimpleVectorQuantizer."""

    def __init__(self, clip_model: "oclip.CLIP", d_model: int, keyword_num: int = 8, nhead: int = 1, vq_temp: float = 0.1,
                 sot_token: int = 1, eot_token: int = 2):
        super().__init__()
        self.clip_model = [clip_model]  # not registered: the reference registers it twice, the state dict is compared per branch
        self.keyword_num = keyword_num
        text_dim = clip_model.token_embedding.weight.shape[1]
        self.cls = nn.Parameter(torch.randn(1, keyword_num, d_model))
        self.self_att = AttentionAndNorm(d_model=d_model, nhead=nhead)
        self.linear_proj = nn.Linear(d_model, text_dim)
        emb = clip_model.token_embedding.weight.detach()
        self.bn_layer = KwBatchNorm(keyword_num, text_dim, emb.mean(0), emb.std(0))
        self.vq_temp = vq_temp
        self.sot_token, self.eot_token = sot_token, eot_token

    def forward(self, audio_feat, audio_len, training=True, collect=None, force_idx=None, attn_mask_rows=None):
        B, T = audio_feat.shape[:2]
        K = self.keyword_num
        src = torch.cat([self.cls.expand(B, K, -1), audio_feat], 1)                     # kwClip.py:870-872
        kpm = keypadding_mask(T + K, audio_len + K)                                        # :874-876
        kw = self.self_att(src, kpm, attn_mask_rows)[:, :K]                                # :878-882
        kw = self.linear_proj(kw)                                                          # :884
        kw = self.bn_layer(kw)                                                             # :886-887
        emb = self.clip_model[0].token_embedding.weight
        cos = torch.stack([F.cosine_similarity(kw[:, i, :].unsqueeze(-1), emb.t().unsqueeze(0), dim=1) for i in range(K)], 1)  # :890-900
        vq = simple_vector_quantizer(cos, self.vq_temp, training, force_idx=force_idx)     # :909
        keywords = vq["subword_prob"] @ emb                                                # :911
        feat = self.clip_model[0].encode_keywords(keywords, K, self.sot_token, self.eot_token)  # :914
        if collect is not None:
            collect.update(kw_bn=kw, cos=cos, keywords=keywords)
        return feat, vq, keywords


def masked_contrastive_loss(feat_a, feat_b, index=None, temperature=1.0 / 0.07, margin=0.0, dcl=False,
                            a2b=True, b2a=True, return_logits=False):
    """losses.py:185-245.  ``temperature`` is the multiplier (1/0.07 fixed, or exp(param) when learnable)."""
    B = feat_a.shape[0]
    eye = torch.eye(B, dtype=torch.bool)
    neg = (index[:, None] != index[None, :]) if index is not None else ~eye
    if not dcl:
        neg = neg | eye
    logits = feat_a @ feat_b.t() * temperature
    if margin > 0.0:
        logits = logits - margin * eye.to(logits.dtype)
    pos = logits.diagonal()
    e = logits.exp() * neg.to(logits.dtype)
    loss = 0
    if a2b:
        loss = loss + (-pos + torch.log(e.sum(1))).mean()
    if b2a:
        loss = loss + (-pos + torch.log(e.sum(0))).mean()
    if a2b and b2a:
        loss = loss / 2
    return (loss, logits) if return_logits else loss


def mutual_retrieval(score_a, score_b, ab_answers, ba_answers, recall_at):
    """retrieval.py:6-121 without the python row loops: recall@k in percent, both directions + mean."""
    def one(score, cand_ids, answers):
        order = torch.argsort(score, dim=1, descending=True)
        hit = cand_ids[order] == answers[:, None]
        return {f"recall@{k}": 100.0 * hit[:, :min(k, hit.shape[1])].any(1).float().sum().item() / hit.shape[0]
                for k in recall_at}
    ab = one(score_a, ba_answers, ab_answers)
    ba = one(score_b, ab_answers, ba_answers)
    return ab, ba, {k: (ab[k] + ba[k]) / 2.0 for k in ab}


def linear_warmup_decay(step: int, base_lr: float, warmup: int, max_step: int, final_lr: float) -> float:
    """scheduler.py:22-38: LR multiplier at LambdaLR step ``step``."""
    if step < warmup:
        return (step + 1) / warmup
    return 1.0 - (1.0 - final_lr / base_lr) * (step + 1 - warmup) / (max_step - warmup)


class _AudioEncoder(nn.Module):
    def __init__(self, cfg: ohubert.HubertCfg):
        super().__init__()
        self.encoder = ohubert.HubertModel(cfg)
        self.weightedsum_layer = nn.Module()
        self.weightedsum_layer.weights = nn.Parameter(torch.zeros(cfg.layers + 1))


class _Clip(nn.Module):
    def __init__(self, cfg: oclip.ClipCfg):
        super().__init__()
        self.model = oclip.CLIP(cfg)


class _Criterion(nn.Module):
    def __init__(self, temperature=0.07, temperature_trainable=False, **_):
        super().__init__()
        self.trainable = temperature_trainable
        if temperature_trainable:
            self.temperature = nn.Parameter(torch.ones([]) * math.log(1 / temperature))
        else:
            self.temperature = 1 / temperature
        eye = torch.eye(256, dtype=torch.bool)  # buffers kept for state-dict key parity (losses.py:165-168)
        self.register_buffer("eye_mat", eye)
        self.register_buffer("neg_eye_mat", ~eye)
        self.register_buffer("eye_mat_fl", eye.float())

    def multiplier(self):
        return self.temperature.exp() if self.trainable else self.temperature


class SpeechClipOracle(nn.Module):
    """Parallel SpeechCLIP (KWClip_GeneralTransformer with parallel_objective_weight > 0), eval-mode arithmetic."""

    def __init__(self, hubert_cfg: ohubert.HubertCfg, clip_cfg: oclip.ClipCfg, branch_args: Optional[dict],
                 loss_args: Optional[dict] = None, normalize_hiddenstates: bool = False, cascaded_args: Optional[dict] = None):
        super().__init__()
        self.audio_encoder = _AudioEncoder(hubert_cfg)
        self.clip = _Clip(clip_cfg)
        self.criterion = _Criterion(**(loss_args or {}))
        self.parallel_branch = self.cascaded_branch = None
        if branch_args is not None:
            self.parallel_branch = ParallelBranch(hubert_cfg.embed_dim, clip_cfg.t_width,
                                                  **{k: v for k, v in branch_args.items() if k != "d_model"})
        if cascaded_args is not None:   # spchclp_c.yaml: cascaded_objective_weight = 1 (kwClip.py:1125-1137)
            self.cascaded_branch = CascadedBranch(self.clip.model, hubert_cfg.embed_dim, **cascaded_args)
        self.normalize_hiddenstates = normalize_hiddenstates
        self.max_audio_len = 102400

    # speech_encoder_plus.py:520-634 (eval mode: no random crop)
    def forward_audio(self, wavs: List[torch.Tensor], return_hidden_states=False, collect=None):
        enc = self.audio_encoder.encoder
        padded, pad_mask = ohubert.preprocess_input(wavs, enc.cfg.normalize_wav)
        out = enc.custom_forward(padded, pad_mask, collect)
        hs = out["layer_results"]
        n_frames = hs[-1].shape[1]
        feat_len = ohubert.feat_lengths([len(w) for w in wavs], n_frames)
        feat = weighted_sum(self.audio_encoder.weightedsum_layer.weights, hs, self.normalize_hiddenstates)
        return (feat, feat_len, tuple(hs)) if return_hidden_states else (feat, feat_len)

    def forward(self, wavs: List[torch.Tensor], image: torch.Tensor, ids: torch.Tensor):
        """kwClip.py:1385-1478 → dict of L2-normalised features."""
        audio_feat, audio_len = self.forward_audio(wavs)
        image_feat = self.clip.model.encode_image(image)
        image_feat = image_feat / image_feat.norm(dim=-1, keepdim=True)
        out = {"id": ids, "image_feat": image_feat, "audio_feat": audio_feat, "audio_len": audio_len}
        if self.cascaded_branch is not None:
            collect = {}
            c, vq, keywords = self.cascaded_branch(audio_feat, audio_len, training=self.cascaded_training, collect=collect,
                                                   force_idx=self.force_idx)
            out.update(cascaded_audio_feat=c / c.norm(dim=-1, keepdim=True), vq_results=vq, keywords=keywords, cascaded_collect=collect)
        if self.parallel_branch is not None:
            p = self.parallel_branch(audio_feat, audio_len)
            out["parallel_audio_feat"] = p / p.norm(dim=-1, keepdim=True)
        return out

    force_idx = None           # test hook, see simple_vector_quantizer
    cascaded_training = True   # straight-through softmax path of the quantiser (module.training in the reference)

    def compute_loss(self, feats: dict, return_logits=False):
        """kwClip.py:1248-1297 with the active objective's weight = 1 (cascaded XOR parallel in every shipped config)."""
        key = "parallel_audio_feat" if self.parallel_branch is not None else "cascaded_audio_feat"
        return masked_contrastive_loss(feats[key].float(), feats["image_feat"].float(), feats["id"],
                                       self.criterion.multiplier(), return_logits=return_logits)

    def encode_speech(self, wavs):
        audio_feat, audio_len = self.forward_audio(wavs)
        p = self.parallel_branch(audio_feat, audio_len)
        return {"parallel_audio_feat": p / p.norm(dim=-1, keepdim=True)}

    def feature_extractor_s3prl(self, wavs):
        """kwClip.py:1214-1246."""
        audio_feat, audio_len, hs = self.forward_audio(wavs, return_hidden_states=True)
        hs = hs + tuple(self.parallel_branch.extract_hidden_states(audio_feat, audio_len)[1:])
        return hs[-1], hs
