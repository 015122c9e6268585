"""Reference: avssl/util/data_utils.py:4-20.  Kept for API parity (host-side helper; the CUDA attention kernels take
valid-key counts directly — see ``scb_lengths_to_i32`` — so the hot path never materialises this mask)."""
import torch


def get_keypadding_mask(max_length: int, data_lens: torch.Tensor) -> torch.Tensor:
    """bool [bsz, max_length], True = padding (position >= length)."""
    positions = torch.arange(max_length, device=data_lens.device)
    return positions.unsqueeze(0) >= data_lens.unsqueeze(1)
