#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r1b_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/r1b_gemm_shapes_base.csv > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
kill $SMI
cut -c1-300 gpurun_out/r1b_bench_n1.json; tail -2 gpurun_out/r1b_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_n1_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r1b_bench_n1_reference_arm.json
timeout 900 python bench.py --config large --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r1b_gemm_shapes_large.csv > gpurun_out/r1b_bench_n1_large.json 2> gpurun_out/r1b_large.err; cut -c1-300 gpurun_out/r1b_bench_n1_large.json; tail -2 gpurun_out/r1b_large.err
timeout 900 python bench.py --config cascaded --steps 5 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/r1b_gemm_shapes_cascaded.csv > gpurun_out/r1b_bench_n1_cascaded.json 2> gpurun_out/r1b_casc.err; cut -c1-300 gpurun_out/r1b_bench_n1_cascaded.json; tail -2 gpurun_out/r1b_casc.err
