"""Tile-shape sweep of scb_gemm on the small-M shapes of strong scaling (CLIP ViT / HuBERT at 32-128 pairs per GPU).
Each candidate runs in a fresh process because the tile override is read once from the environment."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {  # name: (N, K, out fp32 + residual?, act)
    "qkv": (2304, 768, False, 0), "out": (768, 768, True, 0), "fc1": (3072, 768, False, 2), "fc2": (768, 3072, True, 0),
}
CANDS = [("auto", {}), ("bn64", {"SCB_GEMM_FORCE_BN": "64", "SCB_GEMM_FORCE_2CTA": "0"}),
         ("bn128", {"SCB_GEMM_FORCE_BN": "128", "SCB_GEMM_FORCE_2CTA": "0"}),
         ("bn256", {"SCB_GEMM_FORCE_BN": "256", "SCB_GEMM_FORCE_2CTA": "0"}), ("pair", {"SCB_GEMM_FORCE_2CTA": "1"})]


def child(ms):
    import torch
    from speechclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for M in ms:
        for name, (N, K, res, act) in SHAPES.items():
            a = [torch.randn(M, K, device="cuda", generator=g).half() for _ in range(4)]
            w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
            b = torch.randn(N, device="cuda", generator=g)
            r = torch.randn(M, N, device="cuda", generator=g) if res else None
            out = torch.empty(M, N, device="cuda", dtype=torch.float32 if res else torch.float16)
            for i in range(3):
                ops.gemm(a[i % 4], w, bias=b, act=act, residual=r, out=out)
            torch.cuda.synchronize()
            # 20 launches captured into one CUDA graph: the replay has no python / ctypes time between kernels (a single launch
            # from python costs ~20 us of host time, more than these kernels run)
            reps = 20
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(reps):
                    ops.gemm(a[i % 4], w, bias=b, act=act, residual=r, out=out)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            print(f"{os.environ.get('SWEEP_TAG'):6s} M{M:6d} {name:4s} N{N} K{K}: {us:7.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    if os.environ.get("SWEEP_TAG"):
        child([int(x) for x in sys.argv[1:]])
    else:
        ms = sys.argv[1:] or ["1600", "3200", "6400", "10208", "20416"]
        for tag, env in CANDS:
            subprocess.run([sys.executable, __file__] + ms, env={**os.environ, **env, "SWEEP_TAG": tag})
