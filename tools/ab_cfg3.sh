#!/bin/bash
# A/B of library variants on the instanced scene (configs[2], 8 spp): Mrays/s, trace / shade ms, k_trace fraction of the HBM roofline
for v in "$@"; do
  PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so python tools/bench_configs.py --only 3 --spp3 8 --steps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['mrays_per_s'],1), 'Mrays/s  step', round(d['ms_per_step'],1), 'ms  trace', round(d['trace_ms'],1), 'shade', round(d['shade_ms'],1), 'frac', round(d['trace_frac_of_hbm_peak'],3))"
done
