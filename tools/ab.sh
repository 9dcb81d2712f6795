#!/bin/bash
# A/B of library variants on the 8-spp bench workload: prints value, ms/step, trace ms, shade ms
for v in "$@"; do
  PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so python bench.py --steps 3 --warmup 3 --spp ${SPP:-8} --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['trace_ms_per_step'],2), 'shade', round(r['shade_ms_per_step'],2), 'frac', round(r['frac'],3))"
done
