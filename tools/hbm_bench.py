"""Time the HBM-bound kernels of the step at the headline shapes (256 pairs, T = 319) against their algorithmic bytes:

    python tools/hbm_bench.py [name ...]      # names: ln_hubert ln_vit wsum_fwd wsum_bwd conv0 (default: all)

CUDA events around `reps` back-to-back launches over rotating buffers larger than the 126 MB L2; peak = MEASURED_PEAKS.json.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechclip_b200 import ops  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6540.8
B, T, D, L = 256, 319, 768, 13
M = B * T
dev = "cuda"
H = torch.float16


def timeit(fn, reps=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:34s} {us:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / PEAK:5.2f} of {PEAK:.0f}")


def ln(rows, name, D=D):
    xs = [torch.randn(rows, D, device=dev) for _ in range(3 if rows > 50000 else 8)]
    ys = [torch.empty(rows, D, device=dev, dtype=H) for _ in xs]
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    us = timeit(lambda i: ops.layernorm(xs[i % len(xs)], g, b, y16=ys[i % len(xs)], rows=rows, d=D), reps=24)
    report(name + f" fp32->fp16 [{rows}x{D}]", us, rows * D * 6)


def wsum_large(which):
    Ll, Bl, Dl = 25, 64, 1024   # HuBERT-large, 64 pairs per GPU (BASELINE.json configs[3]); fp32 hidden states, normalised
    Ml = Bl * T
    h = torch.randn(Ll, Ml, Dl, device=dev)
    w = torch.zeros(Ll, device=dev)
    if which == "fwd":
        out = torch.empty(Ml, Dl, device=dev)
        us = timeit(lambda i: ops.weighted_sum(h, w, True, out32=out), reps=6)
    else:
        dout = torch.randn(Bl, T, Dl, device=dev)
        gw = torch.zeros(Ll, device=dev)
        scratch = torch.empty(64, device=dev)
        us = timeit(lambda i: ops.weighted_sum_bwd(h, w, True, dout, T, dout.stride(0), 0, scratch, gw, 1.0), reps=6)
    report(f"weighted_sum_{which} norm fp32 [{Ll}x{Ml}x{Dl}]", us, Ll * Ml * Dl * 4 + Ml * Dl * 4)


def wsum(which):
    h = (torch.randn(L, M, D, device=dev, dtype=H))
    w = torch.zeros(L, device=dev)
    if which == "fwd":
        out = torch.empty(M, D, device=dev)
        us = timeit(lambda i: ops.weighted_sum(h, w, False, out32=out), reps=6)
        report(f"weighted_sum_fwd fp16 [{L}x{M}x{D}]", us, L * M * D * 2 + M * D * 4)
    else:
        dout = torch.randn(B, T, D, device=dev)
        gw = torch.zeros(L, device=dev)
        scratch = torch.empty(64, device=dev)
        us = timeit(lambda i: ops.weighted_sum_bwd(h, w, False, dout, T, dout.stride(0), 0, scratch, gw, 1.0), reps=6)
        report(f"weighted_sum_bwd fp16 [{L}x{M}x{D}]", us, L * M * D * 2 + M * D * 4)


def conv0():
    Tw = 102400
    Tf = (Tw - 10) // 5 + 1
    wav = 0.1 * torch.randn(B, Tw, device=dev)
    w, cb = 0.3 * torch.randn(512, 10, device=dev), torch.zeros(512, device=dev)
    g, b = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    out = torch.empty(B * Tf * 512 + 4096, device=dev, dtype=H)
    scratch = torch.empty(ops.conv0_scratch_bytes(B), device=dev, dtype=torch.uint8)
    us = timeit(lambda i: ops.conv0_groupnorm_gelu(wav, Tw, w, cb, g, b, 1e-5, out, Tf * 512, scratch), reps=4)
    report(f"conv0+GroupNorm+GELU [{B}x{Tf}x512] fp16", us, B * Tf * 512 * 2 + 2 * B * Tw * 4)


def conv0_large():
    Tw = 102400
    Tf = (Tw - 10) // 5 + 1
    Bl = 64
    wav = 0.1 * torch.randn(Bl, Tw, device=dev)
    w, cb = 0.3 * torch.randn(512, 10, device=dev), 0.1 * torch.randn(512, device=dev)
    g, b = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    out = torch.empty(Bl * Tf * 512 + 4096, device=dev, dtype=H)
    scratch = torch.empty(ops.conv0_scratch_bytes(Bl), device=dev, dtype=torch.uint8)
    us = timeit(lambda i: ops.conv0_layernorm_gelu(wav, Tw, w, cb, g, b, 1e-5, out, Tf * 512, scratch), reps=4)
    report(f"conv0+LayerNorm+GELU [{Bl}x{Tf}x512] fp16", us, Bl * Tf * 512 * 2 + Bl * Tw * 4)


def ln_conv_large():
    Bl, rows_per = 64, 10239
    rows = Bl * rows_per
    xs = [torch.randn(rows, 512, device=dev).half() for _ in range(2)]
    g, b = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    us = timeit(lambda i: ops.layernorm(xs[i % 2], g, b, y16=xs[i % 2], rows=rows, d=512, act=ops.ACT_GELU), reps=8)
    report(f"layernorm+GELU in place fp16 [{rows}x512]", us, rows * 512 * 4)


ALL = {"conv0_large": conv0_large, "ln_conv_large": ln_conv_large,"ln_hubert": lambda: ln(M, "layernorm (HuBERT)"), "ln_vit": lambda: ln(B * 50, "layernorm (ViT-B/32)"),
       "ln_large": lambda: ln(M, "layernorm (HuBERT-large)", 1024), "ln_vit_l": lambda: ln(64 * 257, "layernorm (ViT-L/14, 64)", 1024),
       "wsum_large_fwd": lambda: wsum_large("fwd"), "wsum_large_bwd": lambda: wsum_large("bwd"),
       "wsum_fwd": lambda: wsum("fwd"), "wsum_bwd": lambda: wsum("bwd"), "conv0": conv0}
for n in (sys.argv[1:] or list(ALL)):
    ALL[n]()
