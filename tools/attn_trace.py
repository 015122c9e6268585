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
names = ["start", "s_full0", "max0", "pvwait", "ldO", "drain", "p_full0", "s_full1", "max1", "p_full1", "end", "M:PV0", "M:S0", "M:PV1", "M:S1"]
slots = [7, 0, 1, 4, 14, 2, 3, 5, 6, 8, 9, 10, 11, 12, 13]
print("tile  " + " ".join(f"{n:>8s}" for n in names))
for i in range(12, 18):
    print(f"{i:4d}  " + " ".join(f"{t[i][j] - base:8d}" for j in slots))
print("quarter 1 warp:")
for i in range(12, 18):
    print(f"{i:4d}  " + " ".join(f"{t[i + 32][j] - base:8d}" for j in slots[:11]))
