"""HuBERT speech encoder (reference: avssl/module/speech_encoder_plus.py:319-634, FairseqSpeechEncoder_Hubert).

The reference loads fairseq's HubertModel and monkey-patches it to return every layer's output; here the same parameters
(held under fairseq's state-dict names) are compiled into a ``speechclip_b200.engine.HubertPlan`` and the forward is a
sequence of sm_100a kernels.  Host-side per-sample python loops and ``.item()`` syncs of the reference (:539-552,602-611)
are replaced by one bookkeeping kernel (``scb_frame_lengths``) + one crop/pad kernel (``scb_wav_prepare``).
"""
import logging
import os
import pickle
import types
from typing import List, Tuple, Union

import torch
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from speechclip_b200 import engine, ops
from speechclip_b200.engine import HubertPlan, conv_out_len
from speechclip_b200.functional import workspace
from speechclip_b200.init import seeded_init_
from speechclip_b200.params import HubertArch, ParamTree, hubert_param_shapes, restoring

from ..util import freeze_model
from .weighted_sum import WeightedSumLayer

logger = logging.getLogger(__name__)

FEAT_SELECT_IDX_WEIGHTED_SUM_MODE = "weighted_sum"


class _MissingClass(dict):
    """Stand-in for classes of packages that are not installed (omegaconf / fairseq / argparse namespaces inside the ``cfg`` and
    ``task_state`` entries of a fairseq checkpoint): only the ``model`` state dict of such a file is read.  It accepts whatever
    the pickle stream does to the object it replaces (constructor arguments, state, dict items, list appends)."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def __setstate__(self, state):
        self["__state__"] = state

    def append(self, item):
        self.setdefault("__items__", []).append(item)

    def extend(self, items):
        self.setdefault("__items__", []).extend(items)

    add = append


class _LenientUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ModuleNotFoundError, AttributeError):
            return _MissingClass


def load_checkpoint_lenient(path: str):
    """``torch.load`` that survives pickled objects of absent packages (fairseq checkpoints carry omegaconf configs)."""
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except (ModuleNotFoundError, AttributeError):
        lenient = types.ModuleType("scb_lenient_pickle")
        lenient.Unpickler = _LenientUnpickler
        lenient.load = lambda f, **kw: _LenientUnpickler(f, **kw).load()
        lenient.__name__ = "pickle"
        return torch.load(path, map_location="cpu", weights_only=False, pickle_module=lenient)


class FairseqSpeechEncoder_Hubert(nn.Module):
    MODEL2URL = {
        "hubert": "https://dl.fbaipublicfiles.com/hubert/hubert_base_ls960.pt",
        "hubert_base": "https://dl.fbaipublicfiles.com/hubert/hubert_base_ls960.pt",
        "hubert_large_ll60k": "https://dl.fbaipublicfiles.com/hubert/hubert_large_ll60k.pt",
    }
    MODEL_DOWNSAMPLE_RATE = {"hubert": 320, "hubert_base": 320, "hubert_large_ll60k": 320}
    TEST_ARCHS = ("tiny", "tiny_large")  # structure-preserving miniatures for parity tests

    def __init__(self, name: str, pretrained: bool = False, trainable: bool = False, device: str = "cpu",
                 feat_select_idx: Union[str, list] = "all", layer_drop: Union[str, float] = 0.0, max_audio_len: int = -1,
                 reinit_layers: List[int] = [], unfreeze_layers: List[int] = [], normalize_hiddenstates: bool = False,
                 normalize_type: str = "s3prl", **kwargs):
        super().__init__()
        assert name in self.MODEL2URL or name in self.TEST_ARCHS, "Model name({}) should be in {}".format(name, self.MODEL2URL.keys())
        self.name = name
        self.pretrained = pretrained
        self.trainable = trainable
        self.feat_select_idx = feat_select_idx
        self.max_audio_len = max_audio_len
        self.reinit_layers = reinit_layers
        self.unfreeze_layers = unfreeze_layers
        self.normalize_hiddenstates = normalize_hiddenstates
        assert normalize_type in ["s3prl", "method1", "method2"], normalize_type
        self.normalize_type = normalize_type
        if trainable or len(reinit_layers) > 0 or len(unfreeze_layers) > 0:
            raise NotImplementedError("a trainable HuBERT is outside the B200 hot path: every shipped config freezes it "
                                      "(config/speechCLIP/**/spchclp_*.yaml: audio_encoder.trainable: false)")
        if normalize_hiddenstates and normalize_type != "s3prl":
            raise NotImplementedError("normalize_type method1/method2 are not used by any shipped config")
        if not (layer_drop == "original" or (isinstance(layer_drop, float) and 0.0 <= layer_drop <= 1.0)):
            raise ValueError(f"layer_drop = {layer_drop} is not supported.")

        self.arch = HubertArch.named(name)
        self.encoder = ParamTree.from_shapes(hubert_param_shapes(self.arch))
        seeded_init_(self.encoder, int(kwargs.get("init_seed", 7122)))
        if pretrained and not restoring():  # load_from_checkpoint: the .ckpt's state_dict fills audio_encoder.encoder.* itself
            ckpt = kwargs.get("ckpt_path") or os.environ.get("SPEECHCLIP_HUBERT_CKPT")
            if not ckpt or not os.path.exists(ckpt):
                raise FileNotFoundError(
                    f"pretrained=True needs the fairseq checkpoint ({self.MODEL2URL.get(name)}) on local disk: pass "
                    "audio_encoder.ckpt_path or set SPEECHCLIP_HUBERT_CKPT (this build has no network access)")
            state = load_checkpoint_lenient(ckpt)
            state = state.get("model", state)
            missing, unexpected = self.encoder.load_state_dict(state, strict=False)
            if missing:
                raise KeyError(f"HuBERT checkpoint misses {missing[:5]}...")
        freeze_model(self.encoder)
        self.encoder.eval()
        self.downsample_rate = self.MODEL_DOWNSAMPLE_RATE.get(name, 320)
        self.upstream_model_hiddenstates_len = self.arch.layers + 1
        self.out_dim = self.arch.embed_dim
        logger.info(f"Loaded HuBERT speech encoder ({name}): out_dim = {self.out_dim}")
        if self.feat_select_idx == FEAT_SELECT_IDX_WEIGHTED_SUM_MODE:
            self.weightedsum_layer = WeightedSumLayer(
                n_weights=self.upstream_model_hiddenstates_len,
                normalize_features=self.normalize_hiddenstates and self.normalize_type == "s3prl")
        self._plan = None
        self._plan_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_plan())

    # ------------------------------------------------------------------------------------------------- plan handling
    def invalidate_plan(self):
        self._plan = None

    def _apply(self, fn, *a, **k):
        self._plan = None  # .to()/.cuda()/.half() move the parameters: rebuild the GEMM-layout copies
        return super()._apply(fn, *a, **k)

    def plan(self, device) -> HubertPlan:
        key = (str(device),)
        if self._plan is None or self._plan_key != key:
            a = self.arch
            self._plan = HubertPlan(self.encoder.state_dict(), device, heads=a.heads, layer_norm_first=a.layer_norm_first,
                                    extractor_layer_norm=a.extractor_layer_norm, pos_groups=a.pos_groups)
            self._plan_key = key
        return self._plan

    def trainable_params(self) -> list:
        if self.feat_select_idx == FEAT_SELECT_IDX_WEIGHTED_SUM_MODE:
            logger.info("Adding weightedsum params")
            return list(self.weightedsum_layer.parameters())
        return []

    # ------------------------------------------------------------------------------------------------- forward
    def _batch(self, wav, wav_len) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (padded fp32 [B, Tmax] on the parameter device, int64 lengths [B] on that device or None = all full)."""
        dev = self.encoder.layer_norm.weight.device
        if dev.type != "cuda":
            raise RuntimeError("FairseqSpeechEncoder_Hubert: move the module to a CUDA device first (no CPU path)")
        if isinstance(wav, torch.Tensor):
            if wav.dim() == 1:
                wav = wav.unsqueeze(0)
            wav = wav.to(device=dev, dtype=torch.float32)
            if isinstance(wav_len, torch.Tensor):
                lens = wav_len.to(device=dev, dtype=torch.int64).contiguous() if wav_len.numel() > 0 else None
            elif len(wav_len) > 0:
                lens = torch.tensor([int(x) for x in wav_len], dtype=torch.int64).to(dev)
            else:
                lens = None
            return wav.contiguous(), lens
        wavs = [w.to(device=dev, dtype=torch.float32) for w in wav]
        lens = torch.tensor([len(w) for w in wavs], dtype=torch.int64).to(dev)
        return pad_sequence(wavs, batch_first=True).contiguous(), lens

    def encode_frozen(self, wav: Union[torch.Tensor, list], wav_len: Union[torch.Tensor, list] = [], slot: int = 0) -> dict:
        """The frozen part of ``forward`` — crop / pad / normalise, frame lengths, the HuBERT tower — on the CURRENT stream;
        returns a handle for ``forward(..., frozen=handle)``.  ``slot`` selects one of the independent sets of staging
        and tower buffers (``functional.workspace``; under CUDA-graph replay each set gets its own graph and output slab): a caller that runs the tower of batch i + 1 before the
        backward pass of batch i (``speechclip_b200.runtime.TowerPipeline``) alternates slots."""
        wav, lens = self._batch(wav, wav_len)
        B, Tmax = wav.shape
        dev = wav.device
        crop = self.training and self.max_audio_len > 0 and Tmax > self.max_audio_len
        tw = self.max_audio_len if crop else Tmax
        T = conv_out_len(tw)
        if T < 1:
            raise ValueError(f"utterances of {tw} samples are shorter than HuBERT's receptive field (400 samples)")
        ws = workspace(dev, slot)
        ints = ws.view("len_ints", (4, B), torch.int32)  # persistent: valid_frames' address is part of the tower's graph signature
        crop_off, crop_len, valid_frames, feat_len32 = ints[0], ints[1], ints[2], ints[3]
        feat_len = torch.empty(B, device=dev, dtype=torch.int64)
        u = torch.rand(B, device=dev) if crop else None  # random crop offset (audio_transforms.py:5-23)
        ops.frame_lengths(lens, B, tw, self.max_audio_len if crop else 0, T, self.downsample_rate, u, crop_off, crop_len, valid_frames,
                          feat_len32, feat_len)
        wav_p = ws.view("wav_prepared", (B, tw), torch.float32)
        stats = ws.view("wav_stats", (2 * B,), torch.float32) if self.arch.normalize_wav else None
        ops.wav_prepare(wav, crop_off, crop_len, tw, self.arch.normalize_wav, stats, wav_p)
        with torch.no_grad():
            hidden, T = self.plan(dev).forward(ws, wav_p, valid_frames if lens is not None else None)
        return {"hidden": hidden, "T": T, "B": B, "feat_len": feat_len}

    def forward(self, wav: Union[torch.Tensor, list] = None, wav_len: Union[torch.Tensor, list] = [],
                feat_select_idx: Union[str, list] = None, return_hidden_states: bool = False, frozen: dict = None) -> tuple:
        if frozen is None:
            frozen = self.encode_frozen(wav, wav_len)
        hidden, T, B, feat_len = frozen["hidden"], frozen["T"], frozen["B"], frozen["feat_len"]
        d = self.out_dim
        slab = hidden.view(hidden.shape[0], B, T, d)
        slab._scb_graph_output = hidden if hasattr(hidden, "_scb_generation") else None
        if feat_select_idx is None:
            feat_select_idx = self.feat_select_idx
        hand_out = feat_select_idx != FEAT_SELECT_IDX_WEIGHTED_SUM_MODE or return_hidden_states
        if hand_out and slab.dtype != torch.float32:
            user = slab.float()  # the post-LN tower keeps fp16 hidden states; callers get fp32 tensors like the reference's
        elif hand_out and engine.GRAPHS:
            user = slab.clone()  # hidden states handed to the caller must survive the next forward (the plan reuses its buffers)
        else:
            user = slab

        states = lambda: tuple(user[i] for i in range(user.shape[0]))
        ret = []
        if feat_select_idx == "all":
            ret.extend([{"last_hidden_state": user[-1], "hidden_states": states()}, feat_len])
        elif feat_select_idx == FEAT_SELECT_IDX_WEIGHTED_SUM_MODE:
            ret.extend([self.weightedsum_layer(slab), feat_len])
        elif isinstance(feat_select_idx, list):
            ret.extend([[user[i] for i in feat_select_idx], feat_len])
        elif feat_select_idx == "last_hidden_state":
            ret.extend([user[-1], feat_len])
        elif feat_select_idx == "hidden_states":
            ret.extend([states(), feat_len])
        else:
            raise KeyError(feat_select_idx)
        if return_hidden_states:
            ret.append(states())
        return tuple(ret)

class S3prlSpeechEncoderPlus(nn.Module):
    """Reference: avssl/module/speech_encoder_plus.py:110-336 (speech encoders loaded through the s3prl hub).  Every shipped
    config uses ``audio_encoder.type: FairseqHubert``; the s3prl package and its hub checkpoints are not available offline and
    the upstreams it can load are arbitrary torch models with no B200 plan.  The class keeps the name and constructor
    signature so ``avssl.module`` / ``kwClip.py:61`` resolve, and fails at construction with the reason."""

    def __init__(self, name: str, pretrained: bool = False, trainable: bool = False, device: str = "cpu",
                 feat_select_idx: Union[str, list] = "all", layer_drop: Union[str, float] = 0.0, max_audio_len: int = -1,
                 reinit_layers: List[int] = [], unfreeze_layers: List[int] = [], **kwargs):
        super().__init__()
        raise NotImplementedError(f"S3prlSpeechEncoderPlus({name!r}): s3prl upstreams are outside the B200 hot path; use "
                                  "audio_encoder.type: FairseqHubert (hubert / hubert_large_ll60k), as every shipped config does")
