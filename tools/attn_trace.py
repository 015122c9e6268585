import ctypes, os, sys
import torch
sys.path.insert(0, os.getcwd())
from speechclip_b200 import ops, lib
B, T, heads = 256, 319, 12
d = heads * 64
qkv = (0.7 * torch.randn(B, T, 3 * d, device="cuda")).half()
out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
kv = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(3):
    ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], out, heads, 0.125, kv, False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 16))()
assert lib.load().scb_debug_attn_trace(buf) == 0
t = [[buf[i * 16 + j] for j in range(16)] for i in range(64)]
base = t[6][7]
names = ["start", "s_full0", "max0", "pvwait", "ldO", "drain", "p_full0", "s_full1", "max1", "p_full1", "end", "M:pf0", "M:PV0", "M:S0", "M:PV1", "M:S1"]
slots = [7, 0, 1, 4, 14, 2, 3, 5, 6, 8, 9, 15, 10, 11, 12, 13]
print("tile  " + " ".join(f"{n:>8s}" for n in names))
for i in range(12, 18):
    print(f"{i:4d}  " + " ".join(f"{t[i][j] - base:8d}" for j in slots))

bw = (ctypes.c_longlong * (32 * 16 * 4))()
assert lib.load().scb_debug_attn_trace_w(bw) == 0
print("per softmax warp (q = warp & 3, part = warp >> 2): clocks after the tile's first s_full0")
for i in (13, 16):
    rows = [[bw[((i * 16) + w) * 4 + k] for k in range(4)] for w in range(16)]
    b0 = min(r[0] for r in rows)
    print(f"tile {i}:  warp   q part   s_full0  p_full0  s_full1  p_full1")
    for w, r in enumerate(rows):
        print(f"          {w:3d} {w & 3:3d} {w >> 2:4d}  " + " ".join(f"{x - b0:8d}" for x in r))
