#!/bin/bash
# GPU-box script: parity tests of the GEMM, shape micro-benchmarks with and without the TMA-store epilogue, then the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/t_gemm.log
cat gpurun_out/t_gemm.log
echo "--- TMA store on"; timeout 300 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gb_tma1.log
echo "--- TMA store off"; SCB_GEMM_TMA_STORE=0 timeout 300 python tools/gemm_bench.py qkv fc1 vit_fc1 plain conv1 conv2 conv5 2>&1 | tee gpurun_out/gb_tma0.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/shapes_b256.csv > gpurun_out/bench_b256.json 2> gpurun_out/bench_b256.err
cut -c1-900 gpurun_out/bench_b256.json; tail -3 gpurun_out/bench_b256.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 32 --dump-profile gpurun_out/shapes_b32.csv > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err
cut -c1-600 gpurun_out/bench_b32.json; tail -3 gpurun_out/bench_b32.err
