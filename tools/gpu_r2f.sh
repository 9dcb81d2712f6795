#!/bin/bash
# Round 2, GPU visit F: bounce loop without host round trips - all -m gpu tests, then A/B against the synchronous loop on
# every config (the small ones are where a bubble per bounce shows), shade variant shA, wave-shape sweep on config 5
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
for c in 1 4 2 3; do CFG=$c STEP=0 STEPS=5 timeout 600 bash tools/ab2.sh presync async 2>&1 | tee -a $OUT/ab_sync_$TAG.txt; done
CFG=3 STEP=8 timeout 300 bash tools/ab2.sh shA 2>&1 | tee -a $OUT/ab_sync_$TAG.txt
CFG=2 STEP=16 timeout 300 bash tools/ab2.sh async shA 2>&1 | tee -a $OUT/ab_sync_$TAG.txt
for ws in "0 0" "4200000 8" "8400000 4" "2100000 8" "3200000 6"; do set -- $ws
  python bench.py --config 5 --step-spp 4 --steps 3 --warmup 3 --no-cpu-baseline --wave-paths $1 --wave-segments $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('cfg5 wave_paths $1 segments $2:', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'launches', d['roofline']['trace_launches_per_step'])" | tee -a $OUT/sweep_cfg5_$TAG.txt
done
