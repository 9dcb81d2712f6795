#!/bin/bash
# Round 2, GPU visit Q: k_ri register caps (launch bounds 1 / 4 / 5 CTAs per SM) on the directlighting / whitted cases
OUT=gpurun_out
mkdir -p $OUT
for v in ri1 ri4 ri5; do
  PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so python tools/bench_integrators.py --no-cpu --only direct 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$v', d['case'], d['mrays_per_s'], 'Mrays/s', d['ms_per_step'], 'ms  trace', d['trace_ms'], 'stage', d['stage_ms'])" | tee -a $OUT/ab_r2q.txt
  PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so python tools/bench_integrators.py --no-cpu --only whitted 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$v', d['case'], d['mrays_per_s'], 'Mrays/s', d['ms_per_step'], 'ms  trace', d['trace_ms'], 'stage', d['stage_ms'])" | tee -a $OUT/ab_r2q.txt
done
