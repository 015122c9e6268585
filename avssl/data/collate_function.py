"""Reference: avssl/data/collate_function.py:7-36 (``collate_general``) — the step in front of the hot path.

``collate_general`` keeps the reference's contract exactly (same keys, ``wav_len`` appended, waveforms zero-padded to the batch
maximum, other tensors stacked, python scalars -> LongTensor); it is host code for DataLoader workers.

``collate_packed`` + ``unpack_on_device`` are the B200-side variant of the same step: the waveforms are packed back to back
(sum of lengths, not B x max length: at Flickr8k's length spread that is ~40 % fewer bytes over PCIe), images may stay uint8 HWC
(4x fewer bytes), and the zero-padding (``scb_pad_rows``) and CLIP's ToTensor + Normalize (``scb_image_normalize``) run on the
device.  ``unpack_on_device`` returns the batch dict ``KWClip_GeneralTransformer.forward`` takes.
"""
from typing import Tuple

import torch
from torch.nn.utils.rnn import pad_sequence

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # openai clip/clip.py _transform
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def collate_general(batch: Tuple) -> dict:
    keys = list(batch[0].keys())
    if "wav" in keys and isinstance(batch[0]["wav"], torch.Tensor):
        keys.append("wav_len")
    out = {k: [] for k in keys}
    for row in batch:
        for k in keys:
            out[k].append(len(row["wav"]) if k == "wav_len" else row[k])
    for k, v in out.items():
        if isinstance(v[0], torch.Tensor):
            out[k] = pad_sequence(v, batch_first=True) if k == "wav" else torch.stack(v, dim=0)
        else:
            out[k] = torch.LongTensor(v)
    return out


def collate_packed(batch: Tuple, pin_memory: bool = False) -> dict:
    """Like ``collate_general`` but ``wav`` is ONE 1-D tensor of the utterances back to back (+ ``wav_len``, ``wav_offset``); a
    uint8 HWC ``image`` stays uint8.  Feed the result (after the host->device copy) to ``unpack_on_device``."""
    out = collate_general([{k: v for k, v in row.items() if k != "wav"} for row in batch]) if len(batch[0]) > 1 else {}
    if "wav" in batch[0]:
        wavs = [row["wav"].reshape(-1).float() for row in batch]
        lens = torch.tensor([w.numel() for w in wavs], dtype=torch.int64)
        out["wav"] = torch.cat(wavs)
        out["wav_len"] = lens
        out["wav_offset"] = torch.cumsum(lens, 0) - lens
    if pin_memory:
        out = {k: v.pin_memory() if isinstance(v, torch.Tensor) else v for k, v in out.items()}
    return out


def unpack_on_device(packed: dict, mean=CLIP_MEAN, std=CLIP_STD) -> dict:
    """Device-resident output of ``collate_packed`` -> the batch dict of ``collate_general`` (padded fp32 ``wav`` [B, Tmax],
    normalised fp32 ``image`` [B, 3, H, W]); padding and normalisation run as kernels."""
    from speechclip_b200 import ops
    out = dict(packed)
    if "wav_offset" in packed:
        lens = packed["wav_len"]
        if not packed["wav"].is_cuda:
            raise RuntimeError("unpack_on_device: CUDA tensors required (no CPU path)")
        tmax = int(packed.get("wav_max_len", 0)) or int(lens.max())   # pass wav_max_len (host int) to avoid the device sync
        out["wav"] = ops.pad_rows(packed["wav"].contiguous(), packed["wav_offset"].contiguous(), lens.contiguous(), tmax)
        out.pop("wav_offset")
        out.pop("wav_max_len", None)
    img = packed.get("image")
    if isinstance(img, torch.Tensor) and img.dtype == torch.uint8:
        out["image"] = ops.image_normalize(img.contiguous(), mean, std)
    return out
