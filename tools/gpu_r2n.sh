#!/bin/bash
# Round 2, GPU visit N: uber / substrate / translucent parity; the headline kernels after the BsdfT refactor; full GPU suite
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "extra_materials or infinite_light or tree_integrators" 2>&1 | tail -15
for c in "2 16" "3 8" "1 16"; do set -- $c
python bench.py --config $1 --step-spp $2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('r2n cfg$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'frac trace/shade/step', round(r['k_trace']['frac'],3), round(r['k_shade']['frac'],3), round(r['step']['frac'],3), 'e2e', round(d['e2e']['value'],1))" | tee -a $OUT/ab_r2n.txt
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu_r2n.txt
