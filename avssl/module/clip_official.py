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

from speechclip_b200 import ops
from speechclip_b200.cascaded import TextTowerPlan, Vocabulary
from speechclip_b200.engine import VitPlan
from speechclip_b200.functional import workspace
from speechclip_b200.init import seeded_init_
from speechclip_b200.params import ClipArch, ParamTree, clip_param_shapes, restoring

logger = logging.getLogger(__name__)

_clip_models = {"RN50", "RN101", "RN50x4", "RN50x16", "RN50x64", "ViT-B/32", "ViT-B/16", "ViT-L/14"}
SOT_TOKEN, EOT_TOKEN = 49406, 49407  # openai BPE ids of <|startoftext|> / <|endoftext|>


def _load_clip_state(path: str) -> dict:
    """State dict of an openai CLIP checkpoint: the released files (``ViT-B-32.pt`` ...) are TorchScript archives with fp16
    weights and three extra scalars (clip.load, clip_official.py:50); a plain ``state_dict`` / ``{"state_dict": ...}`` pickle
    also loads."""
    try:
        return torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        obj = torch.load(path, map_location="cpu", weights_only=False)
        if isinstance(obj, dict) and isinstance(obj.get("state_dict"), dict):
            obj = obj["state_dict"]
        if hasattr(obj, "state_dict"):
            obj = obj.state_dict()
        return obj


class ClipModel(nn.Module):
    def __init__(self, name: str, device: str = "cpu", image_encoder_trainable: bool = False, text_encoder_trainable: bool = False,
                 reduce_subword_embbedding: str = None, **kwargs):
        super().__init__()
        assert name in _clip_models or name in ("tiny", "tiny_c"), name
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
        if ckpt and not restoring():  # load_from_checkpoint: the .ckpt's state_dict fills clip.model.* itself
            state = _load_clip_state(ckpt)
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
            # openai's BPE ids; the test miniatures ("tiny_c", 96-entry vocabulary) keep them as the last two rows
            self.startOfTxt_reduced = self.original2Reduced[min(SOT_TOKEN, self.arch.vocab - 2)]
            self.endOfTxt_reduced = self.original2Reduced[min(EOT_TOKEN, self.arch.vocab - 1)]
        self._vit = None
        self._vit_key = None
        self._text = None
        self._vocab = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_plan())

    def invalidate_plan(self):
        self._vit = self._text = self._vocab = None

    def _apply(self, fn, *a, **k):
        self._vit = self._text = self._vocab = None
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
        """Image files -> preprocessed tensor [B, 3, n_px, n_px] on the model's device (clip_official.py:151-164 with openai's
        ``_transform``: bicubic resize of the short side to n_px, centre crop, RGB, ToTensor, Normalize).  Decode / resize / crop run
        on the host with PIL exactly as in openai CLIP (the same library calls, so the same pixels); ToTensor + Normalize run on the
        device from the uint8 pixels (``scb_image_normalize``)."""
        from PIL import Image
        n_px = self.arch.image_size
        dev = self.model.token_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ClipModel.prep_image: move the module to a CUDA device first (no CPU path)")
        pixels = []
        for p in paths:
            img = Image.open(p) if not isinstance(p, Image.Image) else p
            w, h = img.size
            scale = n_px / min(w, h)   # torchvision Resize(n_px): short side -> n_px, long side int(n_px * long / short)
            nw, nh = (n_px, int(n_px * h / w)) if w <= h else (int(n_px * w / h), n_px)
            img = img.resize((nw, nh), Image.BICUBIC)
            left, top = int(round((nw - n_px) / 2.0)), int(round((nh - n_px) / 2.0))   # torchvision CenterCrop
            img = img.crop((left, top, left + n_px, top + n_px)).convert("RGB")
            pixels.append(torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()))
        from avssl.data.collate_function import CLIP_MEAN, CLIP_STD
        return ops.image_normalize(torch.stack(pixels).to(dev), CLIP_MEAN, CLIP_STD)

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

    def text_plan(self, device) -> TextTowerPlan:
        """Frozen text transformer in GEMM layout (+ fp32 transposed weights for the activation-gradient backward)."""
        key = str(device)
        if self._text is None or self._text[0] != key:
            sd = {k: v for k, v in self.model.state_dict().items() if not k.startswith("visual.")}
            self._text = (key, TextTowerPlan(sd, device, heads=self.arch.t_heads))
        return self._text[1]

    def vocabulary(self, device) -> Vocabulary:
        """The (reduced) token-embedding table prepared for the cosine / VQ kernels."""
        key = str(device)
        if self._vocab is None or self._vocab[0] != key:
            self._vocab = (key, Vocabulary(self.model.token_embedding.weight, device))
        return self._vocab[1]

    def special_tokens(self) -> tuple:
        """(sot, eot) row indices into ``model.token_embedding`` (clip_official.py:236-243)."""
        if self.selected_text_emb_ids is None:
            return (min(SOT_TOKEN, self.arch.vocab - 2), min(EOT_TOKEN, self.arch.vocab - 1))
        return self.startOfTxt_reduced, self.endOfTxt_reduced

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        """Token ids [B, L] (indices into ``model.token_embedding``) -> text features [B, D]; the feature is read at
        ``text.argmax(-1)`` (the [EOT] position) as in openai CLIP.encode_text (clip_official.py:211-218)."""
        if not text.is_cuda:
            raise RuntimeError("ClipModel.encode_text: CUDA tensor required (no CPU path)")
        plan, vocab = self.text_plan(text.device), self.vocabulary(text.device)
        B, L = text.shape
        if L > plan.context:
            raise ValueError(f"text length {L} exceeds the context length {plan.context}")
        text = text.to(torch.int64).contiguous()
        x0 = torch.empty(B, L, plan.d, device=text.device, dtype=torch.float32)
        ops.token_embed(vocab.E, plan.pos, text, x0)
        return plan.forward(workspace(text.device), x0, text.argmax(dim=-1).contiguous(), save=False)[0]

    @torch.no_grad()
    def encode_keywords(self, keywords: torch.Tensor, keyword_num: int) -> torch.Tensor:
        """Keyword embeddings [B, K, W] -> text features [B, D] (clip_official.py:220-268): [SOT], the K keyword vectors,
        [EOT]; the causal transformer is evaluated on those K+2 positions only.  Inference surface (no autograd): training
        differentiates through ``KW_CascadedBranch.forward``, which fuses this with the quantiser."""
        if not isinstance(keywords, torch.Tensor):
            raise TypeError(f"Unknown keywords type {type(keywords)}")
        if not keywords.is_cuda:
            raise RuntimeError("ClipModel.encode_keywords: CUDA tensor required (no CPU path)")
        plan, vocab = self.text_plan(keywords.device), self.vocabulary(keywords.device)
        B, K, W = keywords.shape
        assert K == keyword_num and W == plan.d
        sot, eot = self.special_tokens()
        return self._encode_rows(plan, vocab, keywords.float().contiguous(), sot, eot)

    def _encode_rows(self, plan, vocab, keywords, sot, eot):
        B, K, W = keywords.shape
        dev = keywords.device
        L = K + 2
        x0 = torch.empty(B, L, W, device=dev, dtype=torch.float32)
        tok = torch.zeros(B, L, device=dev, dtype=torch.int64)
        tok[:, 0], tok[:, K + 1] = sot, eot
        ops.token_embed(vocab.E, plan.pos, tok, x0)                       # every row = E[token] + pos
        # rows 1..K: keywords + pos  (residual-add of the positional rows onto the given vectors)
        pos_rows = plan.pos[1:K + 1].reshape(1, K * W).expand(B, K * W).contiguous()
        mid = x0.view(B, L * W)[:, W:(K + 1) * W]
        ops.rows_bias_act(keywords.view(B, K * W), None, pos_rows, K * W, ops.ACT_NONE, None, mid, rows=B, d=K * W, x_ld=K * W, y_ld=L * W)
        return plan.forward(workspace(dev), x0, K + 1, save=False)[0]
