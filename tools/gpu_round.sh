#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_cascaded_gpu.py -x -q 2>&1 | tail -4
for ov in 1 0; do
  SCB_OVERLAP_TOWERS=$ov timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b256_ov$ov.json 2> gpurun_out/bench_b256_ov$ov.err
  cut -c1-330 gpurun_out/bench_b256_ov$ov.json; tail -3 gpurun_out/bench_b256_ov$ov.err
  SCB_OVERLAP_TOWERS=$ov timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 32 > gpurun_out/bench_b32_ov$ov.json 2> gpurun_out/bench_b32_ov$ov.err
  cut -c1-330 gpurun_out/bench_b32_ov$ov.json; tail -3 gpurun_out/bench_b32_ov$ov.err
done
