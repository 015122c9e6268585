cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/r2_smoke.txt; cat gpurun_out/r2_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/r2_gemm_shapes_base.csv > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
kill $SMI
python bench.py --config large --steps 10 --no-cpu-baseline --dump-profile gpurun_out/r2_gemm_shapes_large.csv > gpurun_out/r2_bench_n1_large.json 2>/dev/null
python bench.py --config cascaded --steps 20 --no-cpu-baseline --dump-profile gpurun_out/r2_gemm_shapes_cascaded.csv > gpurun_out/r2_bench_n1_cascaded.json 2>/dev/null
for f in n1 n1_large n1_cascaded; do python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['passes_ms_per_step'], d['roofline']['achieved'], d['roofline']['traffic'], d['clocks']['sm_mhz'])"; done
python tools/gemm_bench.py > gpurun_out/r2_gemm_bench.txt 2>&1; python tools/attn_bench.py >> gpurun_out/r2_gemm_bench.txt 2>&1
python tools/tower_time.py --batch 32 > gpurun_out/r2_tower_time.txt 2>&1; python tools/tower_time.py --batch 256 --reps 8 >> gpurun_out/r2_tower_time.txt 2>&1; cat gpurun_out/r2_tower_time.txt
