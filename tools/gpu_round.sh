#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -4
timeout 900 python tools/gemm_sweep.py 1600 3200 6400 10208 2>&1 | tee gpurun_out/sweep.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/shapes_b256.csv > gpurun_out/bench_b256.json 2> gpurun_out/bench_b256.err
cut -c1-400 gpurun_out/bench_b256.json; tail -3 gpurun_out/bench_b256.err
SCB_HIDDEN_FP32=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b256_h32.json 2> gpurun_out/bench_b256_h32.err
cut -c1-400 gpurun_out/bench_b256_h32.json; tail -3 gpurun_out/bench_b256_h32.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 32 --dump-profile gpurun_out/shapes_b32.csv > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err
cut -c1-400 gpurun_out/bench_b32.json; tail -3 gpurun_out/bench_b32.err
