#!/bin/bash
# Round 2, GPU visit O: shade kernels with the class-aware lobe_type (compare r2n: cfg2 shade 17.55, cfg3 44.43; r2l: 17.21 / 43.56)
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_branches.py -x -q -m gpu 2>&1 | tail -3
for r in 1 2; do
for c in "2 16" "3 8" "1 16"; do set -- $c
python bench.py --config $1 --step-spp $2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('r2o cfg$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'frac trace/shade/step', round(r['k_trace']['frac'],3), round(r['k_shade']['frac'],3), round(r['step']['frac'],3), 'e2e', round(d['e2e']['value'],1))" | tee -a $OUT/ab_r2o.txt
done; done
