#!/bin/bash
# Runs on the GPU box (under gpurun): the evidence set of the round for the bench command.  Outputs land in gpurun_out/ and are
# copied (summarised where large) into profiles/ by hand.  TAG names the set (r1c = end of round 1, r2 = round 2).
set -x
cd "$(dirname "$0")/.."
TAG=${TAG:-r2}
mkdir -p gpurun_out
# 0. parity + smoke
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt
for f in gpurun_out/parity_fullsize_*.json; do [ -f "$f" ] && cp "$f" gpurun_out/${TAG}_$(basename "$f"); done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/${TAG}_smoke.txt
# 1. plain bench runs (never under a profiler), clocks sampled beside the headline run
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/${TAG}_gemm_shapes_base.csv > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_n1_reference_arm.json 2> /dev/null
python bench.py --config large --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/${TAG}_gemm_shapes_large.csv > gpurun_out/${TAG}_bench_n1_large.json 2> /dev/null
python bench.py --config cascaded --steps 20 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/${TAG}_gemm_shapes_cascaded.csv > gpurun_out/${TAG}_bench_n1_cascaded.json 2> /dev/null
# 2. every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graphs off > gpurun_out/${TAG}_launches_bench.json 2> gpurun_out/${TAG}_launches.err
# 3. DRAM traffic of every GEMM launch of the same command
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm --csv \
    --log-file gpurun_out/${TAG}_gemm_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graphs off > /dev/null 2> gpurun_out/${TAG}_gemm_dram.err
# 4. full captures: the heaviest GEMM shapes and the attention kernel
for s in fc1 fc2_16 out16 qkv; do
  ncu --set full --clock-control none --import-source on -k regex:gemm -s 4 -c 1 -f -o gpurun_out/${TAG}_full_$s python tools/gemm_bench.py $s > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_full_attention python tools/attn_bench.py > /dev/null 2>&1
# 5. full captures of the HBM-bound kernels at the headline shapes (tools/hbm_bench.py)
for k in weighted_sum_fwd weighted_sum_bwd conv0_apply; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${TAG}_full_$k python tools/hbm_bench.py wsum_fwd wsum_bwd conv0 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:layernorm_fwd -s 3 -c 1 -f -o gpurun_out/${TAG}_full_layernorm_hubert python tools/hbm_bench.py ln_hubert > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:layernorm_fwd -s 3 -c 1 -f -o gpurun_out/${TAG}_full_layernorm_vit python tools/hbm_bench.py ln_vit > /dev/null 2>&1
python tools/hbm_bench.py > gpurun_out/${TAG}_hbm_bench.txt 2>&1
python tools/tower_time.py --batch 32 > gpurun_out/${TAG}_tower_time.txt 2>&1
python tools/tower_time.py --batch 256 --reps 8 >> gpurun_out/${TAG}_tower_time.txt 2>&1
python tools/gemm_bench.py > gpurun_out/${TAG}_gemm_bench.txt 2>&1
python tools/attn_bench.py >> gpurun_out/${TAG}_gemm_bench.txt 2>&1
cat gpurun_out/${TAG}_pytest_gpu.txt gpurun_out/${TAG}_smoke.txt gpurun_out/${TAG}_gemm_bench.txt
for f in n1 n1_large n1_cascaded; do cut -c1-260 gpurun_out/${TAG}_bench_$f.json; done
