import torch, sys, os
sys.path.insert(0, os.getcwd())
from speechclip_b200 import ops
for (B, T, heads, causal, lens) in [(3, 319, 12, False, [319, 100, 318]), (2, 50, 12, False, None), (2, 77, 8, True, None), (2, 257, 16, False, None), (5, 320, 4, False, [320, 1, 64, 65, 200])]:
    d = heads * 64
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = (0.7 * torch.randn(B, T, 3 * d, device="cuda", generator=g)).half()
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    out = torch.empty(B, T, d, device="cuda", dtype=torch.float16)
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32) if lens else None
    ops.attention(q, k, v, out, heads, 0.125, kv_len, causal=causal)
    qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    if lens:
        pad = torch.arange(T, device="cuda")[None] >= torch.tensor(lens, device="cuda")[:, None]
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    if causal:
        s = s + torch.full((T, T), float("-inf"), device="cuda").triu_(1)
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, d)
    e = (out.float() - ref).abs()
    print(T, heads, causal, "err", e.max().item(), "nan", torch.isnan(out).sum().item(), "bad rows", (e.amax(dim=(0, 2)) > 4e-3).nonzero().view(-1).tolist()[:20])
