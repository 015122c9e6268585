#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 1 0; do
  for b in 32 256; do
    SCB_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch $b 2>gpurun_out/pdl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl $pdl batch $b:', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" || tail -5 gpurun_out/pdl.err
  done
done
