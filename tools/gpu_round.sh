#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r1c_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/r1c_gemm_shapes_base.csv > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err
kill $SMI
python -c "
import json; d=json.load(open('gpurun_out/r1c_bench_n1.json')); print(round(d['ms_per_step'],2), d['e2e'], round(d['roofline']['achieved'],1)); print({k:round(v['ms_per_step'],2) for k,v in list(d['breakdown_ms_per_step'].items())[:8]})"
python bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('default flags:', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['cpu_baseline'])"
