"""Time the grouped positional-conv GEMM (HuBERT pos_conv: k = 128 taps, 16 groups) at 256 utterances:
python tools/posconv_bench.py   (SCB_GEMM_SLAB_MT=1: one M tile per work item; SCB_GEMM_DEBUG=1 prints the tile configuration)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speechclip_b200 import ops  # noqa: E402

for name, B, T, d, G in (("base", 256, 319, 768, 16), ("large", 64, 319, 1024, 16)):
    K, cpg = 128, d // G
    rows_pad = T + K
    g = torch.Generator(device="cuda").manual_seed(0)
    xpad = torch.randn(B, rows_pad, G * 64, device="cuda", generator=g).half()
    w = (torch.randn(G, cpg, K * 64, device="cuda", generator=g) / (K * cpg) ** 0.5).half()
    bias = torch.randn(d, device="cuda", generator=g)
    x = torch.randn(B * T, d, device="cuda", generator=g)
    out = torch.empty(B * T, d, device="cuda")

    def run():
        ops.gemm_raw(a=xpad, a_inner=G * 64, a_rows=rows_pad, a_row_stride=G * 64, a_batch_stride=rows_pad * G * 64, batch=B, m_per_batch=T,
                     w=w, n=cpg, k=K * 64, groups=G, b_group_stride=cpg * K * 64, kb_per_tap=1, tap_row_shift=1, a_group_cols=64, out=out, ldc=d,
                     out_batch_stride=T * d, out_group_cols=cpg, bias=bias, act=ops.ACT_GELU, residual=x, algo_k=K * cpg)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    fl = 2.0 * B * T * d * cpg * K
    print(f"pos-conv {name}: B={B} T={T} d={d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s algorithmic ({fl * (64 / cpg) ** 2 / us / 1e6:7.1f} with the 64-channel padding)")
