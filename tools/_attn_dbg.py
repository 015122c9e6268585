import torch, sys, os
sys.path.insert(0, os.getcwd())
from speechclip_b200 import ops
B, T, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = heads * 64
g = torch.Generator(device="cuda").manual_seed(T)
qkv = (0.7 * torch.randn(B, T, 3 * d, device="cuda", generator=g)).half()
q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
ops.attention(q, k, v, out, heads, 0.125, None, causal=False)
torch.cuda.synchronize()
qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2) for t in (q, k, v))
s = qf @ kf.transpose(-1, -2) * 0.125
ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, d)
print("err", (out.float() - ref).abs().max().item())
