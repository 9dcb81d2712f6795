#!/bin/bash
# Round 2, GPU visit L: infinite light parity; the headline kernels after the env-light / tree-integrator changes (compare
# with profiles/r2i_ab_halton_batch.txt "regs": cfg2 1359.7 Mrays/s trace 20.24 shade 17.06, cfg3 404.4 / 127.22 / 43.04)
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "infinite_light or tree_integrators or ao_integrator" 2>&1 | tail -15
for c in "2 16" "3 8" "1 16"; do set -- $c
python bench.py --config $1 --step-spp $2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('r2l cfg$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'frac trace/shade/step', round(r['k_trace']['frac'],3), round(r['k_shade']['frac'],3), round(r['step']['frac'],3), 'e2e', round(d['e2e']['value'],1))" | tee -a $OUT/ab_r2l.txt
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu_r2l.txt
