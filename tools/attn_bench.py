"""Time scb_attention_fwd on the tower shapes (CUDA events, L2-sized rotation of buffers): python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speechclip_b200 import ops  # noqa: E402

SHAPES = [("hubert-base B256", 256, 319, 12, True), ("hubert-base B32", 32, 319, 12, True), ("hubert-large B64", 64, 319, 16, True),
          ("vit-l/14 B64", 64, 257, 16, False), ("vit-b/32 B256", 256, 50, 12, False)]
for name, B, T, heads, ragged in SHAPES:
    d = heads * 64
    bufs = []
    for i in range(3):
        g = torch.Generator(device="cuda").manual_seed(i)
        bufs.append((0.7 * torch.randn(B, T, 3 * d, device="cuda", generator=g)).half())
    out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
    kv_len = torch.full((B,), T, device="cuda", dtype=torch.int32) if ragged else None
    def run(i):
        qkv = bufs[i % 3]
        ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], out, heads, 0.125, kv_len, False)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 4.0 * B * heads * T * T * 64
    print(f"{name:20s} T={T:4d} heads={heads:2d}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (algorithmic 4 T^2 d)")
