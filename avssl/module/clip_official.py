"""CLIP towers (reference: avssl/module/clip_official.py:26-294, which wraps openai ``clip.load``).

Parameters live under openai CLIP's state-dict names (``clip.model.*`` in a SpeechCLIP checkpoint); ``encode_image`` runs
the VisionTransformer as sm_100a kernels through ``speechclip_b200.engine.VitPlan``.  The openai package, its BPE
vocabulary and its downloaded weights are not available offline: weights come from ``ckpt_path`` (a plain state dict) or
from the deterministic synthetic initialiser, and the tokenizer-dependent helpers raise.
"""
import logging
import os

import numpy as np
import torch
from torch import nn

from speechclip_b200.engine import VitPlan
from speechclip_b200.functional import workspace
from speechclip_b200.init import seeded_init_
from speechclip_b200.params import ClipArch, ParamTree, clip_param_shapes

logger = logging.getLogger(__name__)

_clip_models = {"RN50", "RN101", "RN50x4", "RN50x16", "RN50x64", "ViT-B/32", "ViT-B/16", "ViT-L/14"}
SOT_TOKEN, EOT_TOKEN = 49406, 49407  # openai BPE ids of <|startoftext|> / <|endoftext|>


class ClipModel(nn.Module):
    def __init__(self, name: str, device: str = "cpu", image_encoder_trainable: bool = False, text_encoder_trainable: bool = False,
                 reduce_subword_embbedding: str = None, **kwargs):
        super().__init__()
        assert name in _clip_models or name == "tiny", name
        if name.startswith("RN"):
            raise NotImplementedError("ResNet CLIP towers are outside the B200 hot path (shipped configs use ViT-B/32 and ViT-L/14)")
        if image_encoder_trainable or text_encoder_trainable:
            raise NotImplementedError("trainable CLIP towers are outside the B200 hot path (every shipped config freezes them)")
        self.name = name
        self.device = device
        self.arch = ClipArch.named(name)
        self.model = ParamTree.from_shapes(clip_param_shapes(self.arch))
        seeded_init_(self.model, int(kwargs.get("init_seed", 7122)))
        with torch.no_grad():
            self.model.logit_scale.fill_(float(np.log(1 / 0.07)))
        ckpt = kwargs.get("ckpt_path") or os.environ.get("SPEECHCLIP_CLIP_CKPT")
        if ckpt:
            state = torch.load(ckpt, map_location="cpu")
            self.model.load_state_dict({k: v.float() for k, v in state.items() if k in dict(self.model.named_parameters())}, strict=True)
        self.image_encoder_trainable = image_encoder_trainable
        self.text_encoder_trainable = text_encoder_trainable
        self.out_dim = self.arch.t_width
        self.tokenizer = None  # openai SimpleTokenizer needs its BPE vocabulary file (absent offline)
        self.freeze_models()

        self.selected_text_emb_ids = None
        if reduce_subword_embbedding is not None:
            if not os.path.exists(reduce_subword_embbedding):
                raise FileNotFoundError(f"File not found {reduce_subword_embbedding}")
            data = np.load(reduce_subword_embbedding)
            self.selected_text_emb_ids = data[:, 0]
            dist = data[:, 1]
            self.selected_text_emb_ids_dist = torch.from_numpy(dist / np.sum(dist))
            logger.warning("Reduce text embedding to size of {}".format(len(self.selected_text_emb_ids)))
            self.original_text_emb_weight = self.model.token_embedding.weight
            reduced = self.model.token_embedding.weight.data[torch.from_numpy(self.selected_text_emb_ids).long()]
            holder = ParamTree()
            holder.register_parameter("weight", nn.Parameter(reduced.clone(), requires_grad=False))
            self.model.token_embedding = holder
            self.original2Reduced = {int(old): new for new, old in enumerate(self.selected_text_emb_ids)}
            self.reducedl2Original = {new: int(old) for new, old in enumerate(self.selected_text_emb_ids)}
            self.startOfTxt_reduced = self.original2Reduced[SOT_TOKEN]
            self.endOfTxt_reduced = self.original2Reduced[EOT_TOKEN]
        self._vit = None
        self._vit_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_plan())

    def invalidate_plan(self):
        self._vit = None

    def _apply(self, fn, *a, **k):
        self._vit = None
        return super()._apply(fn, *a, **k)

    def freeze_models(self):
        for p in self.model.parameters():
            p.requires_grad = False

    def trainable_params(self) -> list:
        return []

    def update_device(self, device):
        self.device = device

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.device = self.model.token_embedding.weight.device
        return self

    def prep_image(self, paths: list) -> torch.Tensor:
        raise NotImplementedError("image file preprocessing (PIL + torchvision transforms) is out of scope: pass [B,3,H,W] tensors")

    def prep_text(self, sents: list) -> torch.Tensor:
        raise NotImplementedError("the openai BPE tokenizer is not available offline: pass token tensors")

    def deTokenize(self, sents):
        raise NotImplementedError("the openai BPE tokenizer is not available offline")

    def vit_plan(self, device) -> VitPlan:
        key = str(device)
        if self._vit is None or self._vit_key != key:
            sd = {k: v for k, v in self.model.state_dict().items() if k.startswith("visual.")}
            self._vit = VitPlan(sd, device, heads=self.arch.v_heads)
            self._vit_key = key
        return self._vit

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        """Images [B, 3, H, W] -> features [B, D] (ln_post(x[:,0]) @ proj), fp32."""
        if not image.is_cuda:
            raise RuntimeError("ClipModel.encode_image: CUDA tensor required (no CPU path)")
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != self.arch.image_size or image.shape[3] != self.arch.image_size:
            raise ValueError(f"Incorrect image tensor shape {tuple(image.shape)}")
        return self.vit_plan(image.device).forward(workspace(image.device), image.float().contiguous())

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("CLIP text tower: cascaded-branch row of SURVEY.md §8 (a9), not built yet")

    def encode_keywords(self, keywords: torch.Tensor, keyword_num: int) -> torch.Tensor:
        raise NotImplementedError("CLIP text tower: cascaded-branch row of SURVEY.md §8 (a9), not built yet")
