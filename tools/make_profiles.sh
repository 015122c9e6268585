#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence for the round's bench command.  Outputs land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graphs off > gpurun_out/r1_launches_bench.json 2> gpurun_out/r1_launches.err
# 2. DRAM traffic of every GEMM launch of the same command (one metric pass per kernel)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm --csv \
    --log-file gpurun_out/r1_gemm_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graphs off > /dev/null 2> gpurun_out/r1_gemm_dram.err
# 3. full capture of the top GEMM shapes (fc1 with the GELU epilogue, fc2 with the fp32 residual epilogue)
ncu --set full --clock-control none --import-source on -k regex:gemm -s 4 -c 1 -o gpurun_out/r1_full_fc1 python tools/gemm_bench.py fc1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm -s 4 -c 1 -o gpurun_out/r1_full_fc2 python tools/gemm_bench.py fc2 > /dev/null 2>&1
# 4. clocks during a plain bench run
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r1_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> /dev/null
