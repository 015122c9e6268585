"""SpeechCLIP model (reference: avssl/model/kwClip.py — KWClipBase :49-695, KW_ParallelBranch :1004-1108,
KWClip_GeneralTransformer :1111-1496) for the Parallel configuration, B200-native.

Same LightningModule surface: ``forward(batch) -> (losses, log_metrics, others)``, ``training_step`` returns features,
``training_step_end`` computes the contrastive loss over the GLOBAL batch, ``compute_loss``, ``encode_speech``,
``feature_extractor_s3prl``, ``validation_*``, ``configure_optimizers``.  The reference gets the global batch from
Lightning's single-process DataParallel gather (``strategy: dp``); here it is one process per GPU and
``training_step_end`` all-gathers the pooled embeddings + ids over NCCL (``gather_features``), then every rank evaluates
the masked InfoNCE on the full matrix and back-propagates its own rows.
"""
import logging
import os
from typing import List, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from speechclip_b200 import ops
from speechclip_b200.cascaded import PARAM_ORDER as CASCADED_PARAM_ORDER
from speechclip_b200.cascaded import CascadedHead
from speechclip_b200.functional import CascadedBranchFn, DropoutState, GradArena, L2NormFn, ParallelBranchFn
from speechclip_b200.head import PARAM_ORDER, ParallelHead
from speechclip_b200.optim import FusedAdam

from ..base import OrderedNamespace
from ..module import ClipModel, FairseqSpeechEncoder_Hubert, MLPLayers, losses, mutualRetrieval
from ..module.kw_modules import TransformerModels
from ..module.speechclip_c_modules import vector_quantizers
from ..module.speechclip_c_modules.kw_bn import Kw_BatchNorm
from ..optim import get_scheduler
from .base_model import BaseLightningModel

logger = logging.getLogger(__name__)

__all__ = ["KWClipBase", "KW_CascadedBranch", "KW_ParallelBranch", "KWClip_GeneralTransformer"]

METRIC_REDUCEFN_MAPPING = {
    torch.Tensor: lambda x: torch.mean(x),
    float: lambda x: x,
    int: lambda x: x,
    str: lambda x: x,
}


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    return L2NormFn.apply(x)


def _timed_collective(name: str, fn):
    """Run a torch.distributed call; under bench.py's per-call profiling bracket it with CUDA events like every C-ABI call."""
    prof = ops.PROFILE
    if prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    prof.append((name, 0.0, e0, e1, ""))
    return out


def _all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    out = torch.empty((dist.get_world_size() * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    _timed_collective("nccl_all_gather", lambda: dist.all_gather_into_tensor(out, x))
    return out


class _GatherPacked(torch.autograd.Function):
    """ONE all-gather for the whole feature dict: ids (int64, carried as two fp32 words) and every [B, D_i] fp32 feature matrix
    are packed into one [B, 2 + sum D_i] buffer (NCCL over NVLink on the GPU box; the messages are a few hundred KB, so the call
    count, not the byte count, is what costs).  Backward hands every rank the gradient rows of its own slice — each rank
    evaluates the same global loss, so no reduction is needed here; the per-rank parameter gradients are summed afterwards
    (``KWClipBase.allreduce_gradients``)."""

    @staticmethod
    def forward(ctx, ids, *feats):
        B = feats[0].shape[0]
        widths = [f.shape[1] for f in feats]
        buf = torch.empty(B, 2 + sum(widths), device=feats[0].device, dtype=torch.float32)
        buf[:, :2] = ids.to(torch.int64).contiguous().view(torch.float32).view(B, 2)
        off = 2
        for f, w in zip(feats, widths):
            buf[:, off:off + w] = f
            off += w
        out = _all_gather_rows(buf)
        ctx.rows, ctx.widths = B, widths
        gids = out[:, :2].contiguous().view(torch.int64).view(-1)
        res, off = [gids], 2
        for w in widths:
            res.append(out[:, off:off + w])
            off += w
        ctx.mark_non_differentiable(gids)
        return tuple(res)

    @staticmethod
    def backward(ctx, _gid, *grads):
        r0 = dist.get_rank() * ctx.rows
        return (None,) + tuple(None if g is None else g[r0:r0 + ctx.rows] for g in grads)


def gather_features(feats: dict) -> dict:
    """Per-rank ``{id, image_feat, parallel_audio_feat | cascaded_audio_feat}`` -> the same dict over the global batch (rank-major
    order).  Replaces Lightning's DataParallel gather in front of ``training_step_end`` (kwClip.py:147-167)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return feats
    keys = [k for k, v in feats.items() if k != "id" and isinstance(v, torch.Tensor) and v.dim() == 2 and v.dtype == torch.float32]
    if "id" in feats and isinstance(feats["id"], torch.Tensor) and feats["id"].dim() == 1 and keys:
        packed = _GatherPacked.apply(feats["id"], *[feats[k] for k in keys])
        out = {"id": packed[0], **{k: v for k, v in zip(keys, packed[1:])}}
    else:
        out, keys = {}, []
    for k, v in feats.items():   # anything else (other dtypes / ranks): one call each, or passed through
        if k in out:
            continue
        out[k] = _all_gather_rows(v) if isinstance(v, torch.Tensor) and v.dim() >= 1 else v
    return out


class KWClipBase(BaseLightningModel):
    """Base class for SpeechCLIP (kwClip.py:49-695)."""

    def __init__(self, config: OrderedNamespace):
        super().__init__(config)
        self.audio_encoder_type = config.audio_encoder.type
        if self.audio_encoder_type == "FairseqHubert":
            self.audio_encoder = FairseqSpeechEncoder_Hubert(**config.audio_encoder)
        elif self.audio_encoder_type in ("s3prl", "s3prl_plus"):
            raise NotImplementedError("s3prl speech encoders are outside the B200 hot path (no shipped config selects them)")
        else:
            logger.warning("No audio encoder loaded")
        self.clip = ClipModel(**config.clip)
        if hasattr(self, "audio_encoder"):
            self.audio_embd_dim = self.audio_encoder.out_dim
        self.subword_embd_dim = self.clip.model.token_embedding.weight.size(-1)
        self.recall_at = config.retrieval.recall_at
        self.criterion = getattr(losses, config.cl_loss.type)(**config.cl_loss.args)
        self.log_detokenize_results = config.log_setting.get("log_detokenize_results", True)
        self.keyword_num = self.config.model_settings.cascaded_branch.keyword.number
        self._arena = None

    # ------------------------------------------------------------------------------------------------- arena
    def arena(self) -> GradArena:
        """Flat storage shared by the backward kernels and FusedAdam; rebuilt if ``.to()`` re-homed the parameters."""
        if self._arena is None or not self._arena.intact():
            params = [p for p in self.getTrainableParams() if p.requires_grad]
            if not params or not params[0].is_cuda:
                return None
            self._arena = GradArena(params)
        return self._arena

    def _wire_arena(self):
        fn = self.arena
        if hasattr(self, "audio_encoder") and hasattr(self.audio_encoder, "weightedsum_layer"):
            self.audio_encoder.weightedsum_layer._scb_arena_fn = fn
        self.criterion._scb_arena_fn = fn

    # ------------------------------------------------------------------------------------------------- towers
    def forward_audio(self, wav: Union[torch.Tensor, list], wav_len: Union[torch.Tensor, list] = [],
                      return_hidden_states: bool = False, frozen: dict = None):
        if self.audio_encoder_type in ["s3prl_plus", "FairseqHubert"]:
            return self.audio_encoder(wav, wav_len, return_hidden_states=return_hidden_states, frozen=frozen)
        raise NotImplementedError("Unknown type:{}".format(self.audio_encoder_type))

    def forward_image(self, images: Union[list, torch.Tensor]) -> torch.Tensor:
        if isinstance(images, list):
            image_tensor = self.clip.prep_image(images).to(self.device)
        elif isinstance(images, torch.Tensor):
            if images.dim() != 4 or images.shape[1] != 3:
                raise ValueError(f"Incorrect image tensor shape {images.shape}")
            image_tensor = images
        else:
            raise TypeError(f"Unknown image type {type(images)}")
        return self.clip.encode_image(image_tensor)

    def forward_text(self, sents: Union[list, torch.Tensor]) -> torch.Tensor:
        if isinstance(sents, list):
            text_tensor = self.clip.prep_text(sents).to(self.device)
        elif isinstance(sents, torch.Tensor):
            if sents.dim() != 2:
                raise ValueError(f"Incorrect text tensor shape {sents.shape}")
            text_tensor = sents
        else:
            raise TypeError(f"Unknown text type {type(sents)}")
        return self.clip.encode_text(text_tensor)

    def forward(self, batch: dict) -> tuple:
        raise NotImplementedError()

    def compute_loss(self, input_feats):
        raise NotImplementedError()

    # ------------------------------------------------------------------------------------------------- training hooks
    def training_step(self, batch: dict, batch_idx=None) -> dict:
        losses_, log_metrics = self.forward(batch)[:2]
        return {"loss_feats": losses_, "log_metrics": log_metrics}

    def training_step_end(self, outputs: dict) -> dict:
        if isinstance(outputs, dict):
            if "loss" in outputs:
                return {"loss": torch.mean(outputs["loss"])}
            elif "loss_feats" in outputs and "log_metrics" in outputs:
                losses_ = self.compute_loss(gather_features(outputs["loss_feats"]))
                log_metrics = outputs["log_metrics"]
                result = {
                    **{f"train_{k}": losses_[k] for k in losses_},
                    **{f"train_{k}": METRIC_REDUCEFN_MAPPING[type(log_metrics[k])](log_metrics[k]) for k in log_metrics},
                }
                self.log_dict(result, on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
                return {"loss": losses_["loss"]}
            raise NotImplementedError()
        raise NotImplementedError()

    def on_after_backward(self):
        self.allreduce_gradients()

    def allreduce_gradients(self):
        """Sum the per-rank gradients of the trainable head (one NCCL all-reduce over the flat gradient buffer).

        Every rank back-propagates the same GLOBAL loss through its own rows of the gathered embeddings, so the gradients of the
        branch / weighted-sum parameters are partial sums (-> SUM).  The criterion's own parameters (the learnable temperature of
        ``model_large``) are different: each rank already holds their full gradient, exactly what the reference's single-process
        DataParallel computes once on the master (kwClip.py:184-191).  Only rank 0's copy enters the sum."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if dist.get_rank() != 0:
            for p in self.criterion.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        arena = self.arena()
        if arena is not None:
            for k in (0, 1):
                base = arena.flat_g[k].data_ptr()
                if all(p.grad is not None and p.grad.data_ptr() == base + 4 * off for p, off in zip(arena.params, arena.offsets)):
                    _timed_collective("nccl_all_reduce", lambda: dist.all_reduce(arena.flat_g[k]))
                    return
        for p in (arena.params if arena is not None else self.getTrainableParams()):  # gradients that autograd did not adopt in place
            if p.grad is not None:
                _timed_collective("nccl_all_reduce", lambda: dist.all_reduce(p.grad))

    # ------------------------------------------------------------------------------------------------- validation hooks
    def validation_step(self, batch: dict, batch_idx: int = 0) -> dict:
        losses_, log_metrics, others = self.forward(batch)
        audio_feat = (others["cascaded_audio_feat"] if self.config.retrieval.audio_feat_src == "cascaded"
                      else others["parallel_audio_feat"])
        return_dict = {"id": others["id"], "audio_feat": audio_feat}
        if others.get("image_feat") is not None:
            return_dict["image_feat"] = others["image_feat"]
        return {"loss_feats": losses_, "log_metrics": log_metrics, "others": return_dict}

    def validation_step_end(self, outputs: dict) -> dict:
        assert isinstance(outputs, dict)
        losses_ = self.compute_loss(outputs["loss_feats"])
        log_metrics = outputs["log_metrics"]
        result = {
            **{f"val_{k}": losses_[k] for k in losses_},
            **{f"val_{k}": METRIC_REDUCEFN_MAPPING[type(log_metrics[k])](log_metrics[k]) for k in log_metrics},
        }
        self.log_dict(result, on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        return {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in outputs["others"].items()}

    def validation_epoch_end(self, outputs: list):
        """kwClip.py:468-502: de-duplicate images by id, score = A I^T, recall@k both ways."""
        dev = self.device
        all_ids = torch.cat([x["id"] for x in outputs], dim=0).to(dev)
        all_imgs = torch.cat([x["image_feat"] for x in outputs], dim=0).to(dev)
        all_audio = torch.cat([x["audio_feat"] for x in outputs], dim=0).to(dev).float().contiguous()
        # last occurrence of every id wins, first-seen order is kept (the reference's dict comprehension)
        ids_host = all_ids.cpu().tolist()
        last = {}
        for i, _id in enumerate(ids_host):
            last[_id] = i
        img_ids = list(last.keys())
        all_img_feats = all_imgs[torch.tensor([last[i] for i in img_ids], device=dev)].float().contiguous()
        all_img_ids = torch.tensor(img_ids, dtype=torch.int64, device=dev)
        print("Total #{} images, #{} audio".format(len(all_img_feats), len(all_audio)))
        score_per_audio = torch.empty(all_audio.shape[0], all_img_feats.shape[0], device=dev, dtype=torch.float32)
        ops.sgemm(all_audio, all_img_feats, score_per_audio)
        score_per_image = torch.empty(all_img_feats.shape[0], all_audio.shape[0], device=dev, dtype=torch.float32)
        ops.sgemm(all_img_feats, all_audio, score_per_image)
        return self.reportRetrieval(score_per_A=score_per_audio, score_per_B=score_per_image, AB_answers=all_ids,
                                    BA_answers=all_img_ids)

    def reportRetrieval(self, score_per_A, score_per_B, AB_answers, BA_answers,
                        metadata: dict = {"modality_A_title": "audio", "modality_B_title": "image", "modality_A_logAbbr": "A",
                                          "modality_B_logAbbr": "I"}):
        rAB, rBA, rMean = mutualRetrieval(score_per_A=score_per_A, score_per_B=score_per_B, AB_answers=AB_answers,
                                          BA_answers=BA_answers, recall_at=self.recall_at,
                                          modality_A_title=metadata["modality_A_title"], modality_B_title=metadata["modality_B_title"])
        ab = "{}{}".format(metadata["modality_A_logAbbr"], metadata["modality_B_logAbbr"])
        ba = "{}{}".format(metadata["modality_B_logAbbr"], metadata["modality_A_logAbbr"])
        print(f"val_recall_{ab}", rAB)
        print(f"val_recall_{ba}", rBA)
        print("val_recall_mean", rMean)
        self.log(f"val_recall_{ab}", rAB, sync_dist=True)
        self.log(f"val_recall_{ba}", rBA, sync_dist=True)
        self.log("val_recall_mean", rMean, sync_dist=True)
        if "recall@10" in rMean:
            self.log("val_recall_mean_10", rMean["recall@10"], sync_dist=True)
        return rAB, rBA, rMean

    def processWavs(self, wav) -> Tuple[torch.Tensor, list]:
        return wav, [len(x) for x in wav]

    def feature_extractor_s3prl(self, wav):
        raise NotImplementedError()

    def getTrainableParams(self) -> list:
        my_params = []
        if hasattr(self, "audio_encoder"):
            my_params += self.audio_encoder.trainable_params()
            my_params += list(self.criterion.parameters())
        my_params += self.clip.trainable_params()
        return my_params

    def configure_optimizers(self) -> Tuple[list, list]:
        """kwClip.py:666-694.  ``Adam`` maps to the fused clip+Adam kernel (same update rule); the global-norm clip of the
        Trainer (``trainer.gradient_clip_val``) is applied inside the same pass when ``fuse_grad_clip`` is left on."""
        my_params = self.getTrainableParams()
        name = self.config.audio_encoder.optim.name
        if name != "Adam":
            raise NotImplementedError(f"optimizer {name}: only Adam (every shipped config) has a fused B200 step")
        self._wire_arena()
        clip_val = float(self.config.trainer.get("gradient_clip_val", 0.0) or 0.0) if self.config.get("fuse_grad_clip", True) else 0.0
        audio_optimizer = FusedAdam(my_params, **self.config.audio_encoder.optim.args, max_grad_norm=clip_val, arena_fn=self.arena)
        audio_scheduler = get_scheduler(optimizer=audio_optimizer, **self.config.audio_encoder.scheduler)
        return [audio_optimizer], [{"scheduler": audio_scheduler, "interval": "step"}]


class VQResults(dict):
    """``vq_results`` of the reference (my_vector_quantizer.py:66-164).  ``temp``, ``num_vars`` and ``targets`` are filled
    eagerly; the logging statistics (``code_perplexity``, ``prob_perplexity``, ``ent_per_t``, ``diversity_loss``) and the dense
    one-hot ``subword_prob`` [B, K, V] are computed on first access -- the training step reads none of them."""

    _LAZY_STATS = ("code_perplexity", "prob_perplexity", "ent_per_t", "diversity_loss")

    def __init__(self, vq, cos, idx, stats, bsz, tsz):
        super().__init__(num_vars=cos.shape[1], temp=vq.temperature(), targets=idx.view(bsz, tsz, 1))
        self._src = (vq, cos, idx, stats, bsz, tsz)

    def __missing__(self, key):
        vq, cos, idx, stats, bsz, tsz = self._src
        if key in self._LAZY_STATS:
            for k, v in vq.results(cos, idx, stats, bsz, tsz, produce_targets=False).items():
                self.setdefault(k, v)
            return dict.__getitem__(self, key)
        if key == "subword_prob":
            prob = torch.zeros(bsz * tsz, cos.shape[1], device=cos.device).scatter_(1, idx.view(-1, 1), 1.0).view(bsz, tsz, -1)
            self[key] = prob
            return prob
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._LAZY_STATS or key == "subword_prob"


class KW_CascadedBranch(nn.Module):
    """The cascaded branch (kwClip.py:697-1001): K learned [CLS] queries -> attention block -> projection -> keyword
    BatchNorm -> cosine scores against the CLIP vocabulary -> hard straight-through quantiser -> frozen CLIP text tower."""

    def __init__(self, config: OrderedNamespace, audio_dim: int, text_dim: int, clip: ClipModel) -> None:
        super().__init__()
        self.audio_dim = audio_dim
        self.text_dim = text_dim
        self.clip = clip
        self.config = config
        cb = self.config.model_settings.cascaded_branch
        self.kw_projection_config = cb.keyword.get("kw_projection", None)
        logger.info("Using KW_CascadedBranch")
        self.keyword_num = cb.keyword.number
        self.cls = self._create_cls()
        logger.info("Start init [CLS] {}".format(self.cls.shape))
        assert hasattr(TransformerModels, cb.transformer_type), "transformer structure '{}' not supported".format(cb.transformer_type)
        if cb.transformer_type != "MultiheadAttentionAndNorm":
            raise NotImplementedError("KW_CascadedBranch on B200: transformer_type MultiheadAttentionAndNorm (every shipped cascaded config)")
        logger.info(f"Using {cb.transformer_type} as KW_CascadedBranch")
        self.self_att = getattr(TransformerModels, cb.transformer_type)(**cb.transformer_args)
        if self.kw_projection_config is None:
            logger.info("kw_projection not specified, using single linear layer as default")
            self.linear_proj = nn.Linear(cb.transformer_args.d_model, self.text_dim)
        else:
            raise NotImplementedError("KW_CascadedBranch on B200: keyword.kw_projection (an MLP instead of the single Linear, kwClip.py:757-768) "
                                      "is not wired into the fused keyword path; no shipped config sets it")
        self.vq_type = cb.vq.type
        if not hasattr(vector_quantizers, cb.vq.type):
            raise NotImplementedError("Vq ({}) not implemented".format(cb.vq.type))
        self.vector_quantizer = getattr(vector_quantizers, self.vq_type)(**cb.vq.args)
        if not hasattr(cb.keyword, "batchnorms"):
            raise NotImplementedError("KW_CascadedBranch on B200: keyword.batchnorms is required (every shipped cascaded config sets it)")
        bn = cb.keyword.batchnorms
        emb = self.clip.model.token_embedding.weight
        self.bn_layer = Kw_BatchNorm(kw_num=self.keyword_num, kw_dim=self.text_dim, batchnorm_type=bn.type,
                                     init_bias=torch.mean(emb, dim=0), init_scale=torch.std(emb, dim=0), std_scale=bn.std_scale,
                                     learnable=bn.learnable if hasattr(bn, "learnable") else True,
                                     parallel=bn.parallel if hasattr(bn, "parallel") else False)
        self._scb_arena_fn = None
        self._scb_dropout = DropoutState()

    def _create_cls(self) -> torch.nn.Parameter:
        return torch.nn.Parameter(torch.randn([1, self.keyword_num, self.config.model_settings.cascaded_branch.transformer_args.d_model]))

    def parameters(self, recurse: bool = True):
        """The branch's own parameters.  The reference registers the shared ClipModel as a submodule, so its
        ``parameters()`` also yields the frozen CLIP weights (which the optimiser then ignores); they are skipped here."""
        own = {id(p) for p in self.clip.parameters()}
        return (p for p in super().parameters(recurse) if id(p) not in own)

    def _head(self):
        p = self.self_att.head_params()
        p["cls"] = self.cls
        p["linear_proj.weight"], p["linear_proj.bias"] = self.linear_proj.weight, self.linear_proj.bias
        p["bn_layer.bn_layer.weight"], p["bn_layer.bn_layer.bias"] = self.bn_layer.bn_layer.weight, self.bn_layer.bn_layer.bias
        bn = self.bn_layer.bn_layer
        head = CascadedHead(self.self_att.d_model, self.self_att.nhead, self.keyword_num, self.text_dim, self.self_att.layer_norm_eps,
                            bn.eps, bn.momentum)
        return head, p

    def _kv_len(self, audio_len: torch.Tensor, total_len: int, dev) -> torch.Tensor:
        # get_keypadding_mask(max_length=T+K, data_lens=audio_len+K) as valid-key counts
        kv_len = torch.empty(audio_len.shape[0], device=dev, dtype=torch.int32)
        ops.lengths_to_i32(audio_len.to(device=dev, dtype=torch.int64).contiguous(), self.keyword_num, total_len, kv_len)
        return kv_len

    @torch.no_grad()
    def extract_hidden_states(self, audio_feat: torch.Tensor, audio_len: torch.Tensor) -> Tuple:
        from speechclip_b200.functional import workspace
        head, p = self._head()
        kv_len = self._kv_len(audio_len, audio_feat.size(1) + self.keyword_num, audio_feat.device)
        hidden = head.full_forward(workspace(audio_feat.device), p, audio_feat.float().contiguous(), kv_len)
        return tuple(x[:, self.keyword_num:, ...] for x in hidden)

    def forward(self, audio_feat: torch.Tensor, audio_len: torch.Tensor) -> Tuple[torch.Tensor, dict, torch.Tensor]:
        """-> (audio_feat [B, D] from the CLIP text tower, vq_results, keywords [B, K, W])  (kwClip.py:857-916)."""
        head, p = self._head()
        dev = audio_feat.device
        kv_len = self._kv_len(audio_len, audio_feat.size(1) + self.keyword_num, dev)
        arena = self._scb_arena_fn() if self._scb_arena_fn is not None else None
        bn = self.bn_layer.bn_layer
        sot, eot = self.clip.special_tokens()
        p_drop = float(self.self_att.dropout) if self.training else 0.0   # nn.MultiheadAttention(dropout=...) (TransformerModels.py:110-117)
        rt = dict(vocab=self.clip.vocabulary(dev), text=self.clip.text_plan(dev), bn_buffers=(bn.running_mean, bn.running_var),
                  temp=self.vector_quantizer.temperature(), sot=int(sot), eot=int(eot), training=self.training,
                  need_grad=torch.is_grad_enabled(), drop=(p_drop, self._scb_dropout.advance(dev)) if p_drop > 0 else None)
        feat, keywords, cos, idx, stats = CascadedBranchFn.apply(audio_feat, kv_len, head, arena, rt,
                                                                 *[p[k] for k in CASCADED_PARAM_ORDER])
        if self.training:
            bn.num_batches_tracked += 1
        bsz = audio_feat.size(0)
        return feat, VQResults(self.vector_quantizer, cos, idx, stats, bsz, self.keyword_num), keywords

    def getAttentionMap(self, audio_feat: torch.Tensor, audio_len: torch.Tensor):
        raise NotImplementedError("attention-map / top-k keyword visualisation (kwClip.py:918-1001) needs the BPE tokenizer; out of scope")


class KW_ParallelBranch(nn.Module):
    """The parallel branch (kwClip.py:1004-1108): [CLS] + transformer encoder + projection."""

    def __init__(self, config: OrderedNamespace, audio_dim: int, out_dim: int) -> None:
        super().__init__()
        self.config = config
        self.audio_dim = audio_dim
        self.out_dim = out_dim
        pb = self.config.model_settings.parallel_branch
        self.need_projection = pb.get("need_projection", True)
        assert hasattr(TransformerModels, pb.transformer_type)
        logger.info(f"Using {pb.transformer_type} as KW_ParallelBranch (projection={self.need_projection})")
        self.self_att = getattr(TransformerModels, pb.transformer_type)(**pb.transformer_args)
        self.cls = self._create_cls()
        if self.need_projection:
            self.linear_proj = nn.Linear(self.audio_dim, self.out_dim)
        self._scb_arena_fn = None
        self._scb_dropout = DropoutState()

    def _create_cls(self):
        return torch.nn.Parameter(torch.randn([1, 1, self.config.model_settings.parallel_branch.transformer_args.d_model]))

    def _head(self) -> Tuple[ParallelHead, dict]:
        p = self.self_att.head_params()
        p["cls"] = self.cls
        if self.need_projection:
            p["linear_proj.weight"], p["linear_proj.bias"] = self.linear_proj.weight, self.linear_proj.bias
        else:
            raise NotImplementedError("need_projection: false is not used by any shipped config")
        return ParallelHead(self.self_att.d_model, self.self_att.nhead, self.self_att.layer_norm_eps, True), p

    def _kv_len(self, audio_len: torch.Tensor, total_len: int, dev) -> torch.Tensor:
        # key-padding mask of get_keypadding_mask(max_length=T+1, data_lens=audio_len+1) as valid-key counts
        kv_len = torch.empty(audio_len.shape[0], device=dev, dtype=torch.int32)
        ops.lengths_to_i32(audio_len.to(device=dev, dtype=torch.int64).contiguous(), 1, total_len, kv_len)
        return kv_len

    @torch.no_grad()
    def extract_hidden_states(self, audio_feat: torch.Tensor, audio_len: torch.Tensor) -> Tuple:
        from speechclip_b200.functional import workspace
        head, p = self._head()
        kv_len = self._kv_len(audio_len, audio_feat.size(1) + 1, audio_feat.device)
        _, hidden = head.full_forward(workspace(audio_feat.device), p, audio_feat.float().contiguous(), kv_len)
        return tuple(x[:, 1:, ...] for x in hidden)

    def forward(self, audio_feat: torch.Tensor, audio_len: torch.Tensor) -> torch.Tensor:
        head, p = self._head()
        kv_len = self._kv_len(audio_len, audio_feat.size(1) + 1, audio_feat.device)
        arena = self._scb_arena_fn() if self._scb_arena_fn is not None else None
        p_drop = float(self.self_att.dropout) if self.training else 0.0   # nn.TransformerEncoderLayer(dropout=...) (TransformerModels.py:55-75)
        drop = (p_drop, self._scb_dropout.advance(audio_feat.device)) if p_drop > 0 else None
        return ParallelBranchFn.apply(audio_feat, kv_len, head, arena, drop, *[p[k] for k in PARAM_ORDER])


OVERLAP_TOWERS = os.environ.get("SCB_OVERLAP_TOWERS", "1") != "0"
_SIDE_STREAMS: dict = {}


def _side_stream(device, which: str = "image") -> "torch.cuda.Stream":
    key = (str(device), which)
    if key not in _SIDE_STREAMS:
        # (a higher stream priority for the speech tower was measured: no difference at 32 pairs, 1 % slower at 256)
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class KWClip_GeneralTransformer(KWClipBase):
    """Main class for SpeechCLIP (kwClip.py:1111-1496)."""

    def __init__(self, config: OrderedNamespace) -> None:
        super().__init__(config)
        self.cascaded_branch = None
        self.parallel_branch = None
        if self.config.model_settings.cascaded_objective_weight > 0:
            logger.info("Create Cascaded Branch")
            if self.config.model_settings.cascaded_branch.type == "KW_CascadedBranch":
                self.cascaded_branch = KW_CascadedBranch(config=self.config, audio_dim=self.audio_embd_dim,
                                                         text_dim=self.subword_embd_dim, clip=self.clip)
            else:
                raise NotImplementedError()
        if self.config.model_settings.parallel_objective_weight > 0:
            logger.info("Create Parallel Branch")
            self.parallel_branch = KW_ParallelBranch(config=self.config, audio_dim=self.audio_embd_dim, out_dim=self.subword_embd_dim)
        # optional projection networks (kwClip.py:1147-1187; absent from every shipped YAML)
        ms = self.config.model_settings
        self.img_enc_proj_net = self.p_branch_proj_net = self.c_branch_proj_net = None
        for attr, key in (("img_enc_proj_net", "image_encoder_projection"), ("p_branch_proj_net", "parallel_branch_projection"),
                          ("c_branch_proj_net", "cascaded_branch_projection")):
            spec = ms.get(key, None)
            if spec is not None:
                logger.info(f"{key} dims:{spec.dimensions} droupout:{spec.dropout}")
                setattr(self, attr, MLPLayers(units=spec.dimensions, dropout=spec.dropout))
        self._wire_arena()

    def _wire_arena(self):
        super()._wire_arena()
        if self.parallel_branch is not None:
            self.parallel_branch._scb_arena_fn = self.arena
        if self.cascaded_branch is not None:
            self.cascaded_branch._scb_arena_fn = self.arena

    def getTrainableParams(self) -> list:
        _params = super().getTrainableParams()
        if self.cascaded_branch is not None:
            logger.info("Add cascaded_branch parameters")
            _params += list(self.cascaded_branch.parameters())
        if self.parallel_branch is not None:
            _params += list(self.parallel_branch.parameters())
        if self.img_enc_proj_net is not None:
            logger.info("Add img_enc_proj_net parameters")
            _params += list(self.img_enc_proj_net.parameters())
        if self.p_branch_proj_net is not None:
            logger.info("Add parallel_branch_projection parameters")
            _params += list(self.p_branch_proj_net.parameters())
        return _params

    def feature_extractor_s3prl(self, wav) -> Tuple[torch.Tensor, Tuple]:
        wav, wav_len = self.processWavs(wav)
        audio_feat, audio_len, hidden_states = self.forward_audio(wav, wav_len, return_hidden_states=True)
        assert isinstance(hidden_states, tuple)
        if self.cascaded_branch is not None:
            cascaded_hidden_states = self.cascaded_branch.extract_hidden_states(audio_feat, audio_len)
            assert isinstance(cascaded_hidden_states, tuple)
            hidden_states = hidden_states + tuple(cascaded_hidden_states[1:])
        if self.parallel_branch is not None:
            parallel_hidden_states = self.parallel_branch.extract_hidden_states(audio_feat, audio_len)
            assert isinstance(parallel_hidden_states, tuple)
            hidden_states = hidden_states + tuple(parallel_hidden_states[1:])
        return hidden_states[-1], hidden_states

    def compute_loss(self, input_feats: dict):
        assert isinstance(input_feats, dict)
        assert "id" in input_feats
        assert "cascaded_audio_feat" in input_feats or "parallel_audio_feat" in input_feats
        assert "image_feat" in input_feats
        cascaded_audio_feat = input_feats["cascaded_audio_feat"].float() if "cascaded_audio_feat" in input_feats else None
        parallel_audio_feat = input_feats["parallel_audio_feat"].float() if "parallel_audio_feat" in input_feats else None
        image_feat = input_feats["image_feat"].float()
        id = input_feats["id"]
        losses_ = {"loss": 0}
        # weight 1.0 (every shipped config) adds nothing to the graph; other weights scale through autograd
        for key, w, feat in (("c_cl_loss", self.config.model_settings.cascaded_objective_weight, cascaded_audio_feat),
                             ("p_cl_loss", self.config.model_settings.parallel_objective_weight, parallel_audio_feat)):
            if w > 0:
                losses_[key] = self.criterion(feat_A=feat, feat_B=image_feat, index=id)
                term = losses_[key] if w == 1.0 else w * losses_[key]
                losses_["loss"] = term if isinstance(losses_["loss"], int) else losses_["loss"] + term
        return losses_

    def encode_speech(self, wav) -> dict:
        wav, wav_len = self.processWavs(wav)
        audio_feat, audio_len = self.forward_audio(wav, wav_len)
        cascaded_audio_feat = parallel_audio_feat = vq_results = keywords = None
        if self.cascaded_branch is not None:
            cascaded_audio_feat, vq_results, keywords = self.cascaded_branch(audio_feat=audio_feat, audio_len=audio_len)
            cascaded_audio_feat = l2_normalize(cascaded_audio_feat)
        if self.parallel_branch is not None:
            parallel_audio_feat = self.parallel_branch(audio_feat=audio_feat, audio_len=audio_len)
            if self.p_branch_proj_net is not None:
                parallel_audio_feat = self.p_branch_proj_net(parallel_audio_feat)
            parallel_audio_feat = l2_normalize(parallel_audio_feat)
        return {"cascaded_audio_feat": cascaded_audio_feat, "parallel_audio_feat": parallel_audio_feat, "vq_results": vq_results,
                "keywords": keywords}

    def precompute_towers(self, batch: dict, slot: int = 0, overlap: bool = True) -> dict:
        """Launch the two FROZEN towers for ``batch`` on the tower streams and return at once: ``{"audio": handle, "image_raw":
        tensor, "events": (audio_done, image_done)}``.  Passing the result as ``batch["_scb_towers"]`` makes ``forward`` skip the
        towers and wait for the events instead.  Every shipped configuration freezes both towers, so their outputs for batch
        i + 1 do not depend on the optimizer step of batch i: ``speechclip_b200.runtime.TowerPipeline`` uses this to run them under
        the (latency-bound) head / loss / backward / all-reduce / Adam tail of the previous batch.  ``overlap=False`` runs both
        towers on the current stream instead (per-kernel timing)."""
        image = batch["image"]
        dev = image.device
        self.clip.update_device(self.device)
        if not overlap:
            frozen = self.audio_encoder.encode_frozen(batch["wav"], batch["wav_len"], slot=slot)
            return {"audio": frozen, "image_raw": self.forward_image(image), "events": ()}
        cur = torch.cuda.current_stream(dev)
        sa, si = _side_stream(dev, "audio"), _side_stream(dev, "image")
        sa.wait_stream(cur)   # the inputs are ready (and earlier users of this slot's buffers are done) once `cur` gets here
        si.wait_stream(cur)
        with torch.cuda.stream(si):
            image_raw = self.forward_image(image)
            ev_i = torch.cuda.Event()
            ev_i.record(si)
        with torch.cuda.stream(sa):
            frozen = self.audio_encoder.encode_frozen(batch["wav"], batch["wav_len"], slot=slot)
            ev_a = torch.cuda.Event()
            ev_a.record(sa)
        return {"audio": frozen, "image_raw": image_raw, "events": (ev_a, ev_i)}

    def forward(self, batch) -> tuple:
        wav, wav_len, image, id = batch["wav"], batch["wav_len"], batch["image"], batch["id"]
        self.clip.update_device(self.device)
        pre = batch.get("_scb_towers") if isinstance(batch, dict) else None
        # The two frozen towers are independent until the loss: the image tower runs on a side stream, so its small, latency-bound
        # kernels (50 tokens per image) fill SMs the speech tower's kernels leave idle at their tails (SCB_OVERLAP_TOWERS=0: serial).
        if pre is not None:   # towers launched ahead of time (precompute_towers): wait for them, hand their outputs to this stream
            cur = torch.cuda.current_stream(image.device)
            for ev in pre["events"]:
                cur.wait_event(ev)
            image_raw = pre["image_raw"]
            image_raw.record_stream(cur)
            for t in (pre["audio"]["hidden"], pre["audio"]["feat_len"]):
                t.record_stream(cur)
            audio_feat, audio_len = self.forward_audio(None, frozen=pre["audio"])
        elif OVERLAP_TOWERS and isinstance(image, torch.Tensor) and image.is_cuda:
            cur = torch.cuda.current_stream(image.device)
            side = _side_stream(image.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                image_raw = self.forward_image(image)
            image_raw.record_stream(cur)
            audio_feat, audio_len = self.forward_audio(wav, wav_len)
            cur.wait_stream(side)
        else:
            audio_feat, audio_len = self.forward_audio(wav, wav_len)
            image_raw = self.forward_image(image)
        if self.img_enc_proj_net is not None:
            image_raw = self.img_enc_proj_net(image_raw)
        image_feat = l2_normalize(image_raw)
        losses_ = {"id": id, "image_feat": image_feat}
        log_metrics = {}
        cascaded_audio_feat = parallel_audio_feat = vq_results = keywords = None
        if self.cascaded_branch is not None:
            cascaded_audio_feat, vq_results, keywords = self.cascaded_branch(audio_feat=audio_feat, audio_len=audio_len)
            cascaded_audio_feat = l2_normalize(cascaded_audio_feat)
            losses_["cascaded_audio_feat"] = cascaded_audio_feat
        if self.parallel_branch is not None:
            parallel_audio_feat = self.parallel_branch(audio_feat=audio_feat, audio_len=audio_len)
            if self.p_branch_proj_net is not None:
                parallel_audio_feat = self.p_branch_proj_net(parallel_audio_feat)
            parallel_audio_feat = l2_normalize(parallel_audio_feat)
            losses_["parallel_audio_feat"] = parallel_audio_feat
        if self.config.model_settings.cascaded_objective_weight > 0:
            log_metrics["softmax_temp"] = vq_results["temp"]
        log_metrics.update({"cl_temp": self.criterion.current_temperature})
        return (losses_, log_metrics,
                {"cascaded_audio_feat": cascaded_audio_feat, "parallel_audio_feat": parallel_audio_feat, "image_feat": image_feat, "id": id,
                 "vq_results": vq_results, "keywords": keywords})
