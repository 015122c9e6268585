"""Micro-benchmark of scb_attention_fwd on the HuBERT-base shape (B=256, T=319, 12 heads x 64)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speechclip_b200 import ops

B, T, H, hd = int(os.environ.get("B", 256)), 319, 12, 64
d = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, T, 3 * d, device="cuda", generator=g).half()
out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
kv_len = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(2):
    ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], out, H, hd ** -0.5, kv_len)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], out, H, hd ** -0.5, kv_len)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"attention B{B} T{T} H{H} hd{hd}: {ms*1e3:.1f} us  {4.0*B*H*T*T*hd/ms/1e9:.1f} TF/s")
