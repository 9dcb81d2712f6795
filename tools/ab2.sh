#!/bin/bash
# A/B of library variants through bench.py (round 2 layout): CFG=<config> STEP=<spp per step>; prints value, ms/step, trace / shade ms, fractions
for v in "$@"; do
  PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so python bench.py --config ${CFG:-3} --step-spp ${STEP:-8} --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('$v cfg${CFG:-3}', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'frac trace/shade/step', round(r['k_trace']['frac'],3), round(r['k_shade']['frac'],3), round(r['step']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
done
