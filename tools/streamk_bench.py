"""scb_gemm with and without the stream-K workspace on the shapes whose tile count leaves a partial last wave:
python tools/streamk_bench.py  (CUDA events, 20 launches each, rotating A operands)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speechclip_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, out dtype, act, residual dtype)
    ("hubert B32 qkv", 10208, 2304, 768, torch.float16, 0, None), ("hubert B32 out", 10208, 768, 768, torch.float32, 0, torch.float16),
    ("hubert B32 fc1", 10208, 3072, 768, torch.float16, 1, None), ("hubert B32 fc2", 10208, 768, 3072, torch.float32, 0, torch.float16),
    ("vit B32 qkv", 1600, 2304, 768, torch.float16, 0, None), ("vit B32 out", 1600, 768, 768, torch.float32, 0, torch.float32),
    ("vit B32 fc1", 1600, 3072, 768, torch.float16, 2, None), ("vit B32 fc2", 1600, 768, 3072, torch.float32, 0, torch.float32),
    ("vit B256 qkv", 12800, 2304, 768, torch.float16, 0, None), ("vit B256 out", 12800, 768, 768, torch.float32, 0, torch.float32),
    ("vit B256 fc1", 12800, 3072, 768, torch.float16, 2, None), ("vit B256 fc2", 12800, 768, 3072, torch.float32, 0, torch.float32),
    ("hubert B256 out", 81664, 768, 768, torch.float32, 0, torch.float16), ("hubert B256 fc2", 81664, 768, 3072, torch.float32, 0, torch.float16),
    ("hubert-l B64 out", 20416, 1024, 1024, torch.float32, 0, torch.float32), ("hubert-l B64 fc2", 20416, 1024, 4096, torch.float32, 0, torch.float32),
    ("hubert-l B64 qkv", 20416, 3072, 1024, torch.float16, 0, None), ("hubert-l B64 fc1", 20416, 4096, 1024, torch.float16, 1, None),
]
scratch = torch.zeros(ops.gemm_workspace_bytes(), device="cuda", dtype=torch.uint8)
for name, M, N, K, odt, act, rdt in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    a = [torch.randn(M, K, device="cuda", generator=g).half() for _ in range(3)]
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).to(rdt) if rdt else None
    out = torch.empty(M, N, device="cuda", dtype=odt)
    res = []
    for sk in (None, scratch):
        for i in range(3):
            ops.gemm(a[i % 3], w, bias=b, act=act, residual=r, out=out, scratch=sk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ops.gemm(a[i % 3], w, bias=b, act=act, residual=r, out=out, scratch=sk)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / 20)
    fl = 2.0 * M * N * K
    print(f"{name:18s} M{M:6d} N{N:5d} K{K:5d}: plain {res[0]:7.1f} us {fl / res[0] / 1e6:7.1f} TF/s | stream-K {res[1]:7.1f} us {fl / res[1] / 1e6:7.1f} TF/s", flush=True)
