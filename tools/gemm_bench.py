"""Micro-benchmark of scb_gemm on the shapes of the SpeechCLIP step (CUDA events, L2-cold rotation of operands)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speechclip_b200 import ops

SHAPES = {  # name: (M, N, K, out_dtype, act, bias, residual)
    "qkv": (81664, 2304, 768, torch.float16, 0, True, False),
    "out": (81664, 768, 768, torch.float32, 0, True, True),
    "fc1": (81664, 3072, 768, torch.float16, 1, True, False),
    "fc2": (81664, 768, 3072, torch.float32, 0, True, True),
    "vit_fc1": (12800, 3072, 768, torch.float16, 2, True, False),
    "plain": (81664, 2304, 768, torch.float16, 0, False, False),
    "out16": (81664, 768, 768, torch.float32, 0, True, "f16"),
    "fc2_16": (81664, 768, 3072, torch.float32, 0, True, "f16"),
}
CONVS = {"conv1": (256, 20479, 3), "conv2": (256, 10239, 3), "conv5": (256, 1279, 2)}  # name: (batch, T_in, taps), 512 -> 512 stride 2


def run_conv(name, reps=5):
    B, T, k = CONVS[name]
    T_out = (T - k) // 2 + 1
    g = torch.Generator(device="cuda").manual_seed(0)
    x = [torch.randn(B, T, 512, device="cuda", generator=g).half() for _ in range(2)]
    w = (torch.randn(512, k * 512, device="cuda", generator=g) / (k * 512) ** 0.5).half()
    out = torch.empty(B, T_out, 512, device="cuda", dtype=torch.float16)

    def go(i):
        ops.gemm_raw(a=x[i % 2], a_inner=1024, a_rows=(T + 1) // 2, a_row_stride=1024, a_batch_stride=T * 512, batch=B, m_per_batch=T_out,
                     w=w, n=512, k=k * 512, kb_per_tap=16, tap_row_shift=1, out=out, ldc=512, out_batch_stride=T_out * 512, act=1)
    for i in range(2):
        go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        go(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:8s} B{B} T{T_out} N512 K{k*512}: {ms*1e3:8.1f} us  {2.0*B*T_out*512*k*512/ms/1e9:7.1f} TF/s", flush=True)


def run(name, reps=5):
    M, N, K, odt, act, bias, res = SHAPES[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    a = [torch.randn(M, K, device="cuda", generator=g).half() for _ in range(2)]
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    r = torch.randn(M, N, device="cuda", generator=g) if res else None
    if res == "f16":
        r = r.half()
    out = torch.empty(M, N, device="cuda", dtype=odt)
    for i in range(2):
        ops.gemm(a[i % 2], w, bias=b, act=act, residual=r, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.gemm(a[i % 2], w, bias=b, act=act, residual=r, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:8s} M{M} N{N} K{K}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(SHAPES) + list(CONVS)):
        run_conv(n) if n in CONVS else run(n)
