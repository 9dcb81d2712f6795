#!/bin/bash
# A/B of the wave pipelining (pt_render_opts.wave_segments x wave_paths) on the bench workload; one JSON line per setting.
OUT=gpurun_out
mkdir -p $OUT
for cfg in "$@"; do
  wp=${cfg%%:*}; ws=${cfg##*:}
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --wave-paths $wp --wave-segments $ws > $OUT/pipe_${wp}_${ws}.json 2> $OUT/pipe_${wp}_${ws}.err
  python - <<EOF
import json
l=json.load(open("$OUT/pipe_${wp}_${ws}.json"))
r=l["roofline"]
print("wave_paths %9s segments %s: %7.1f Mrays/s  %6.1f ms/step  trace %6.1f ms shade %6.1f ms  frac %.3f  trace launches %d e2e %.1f" % ("$wp", "$ws", l["value"], l["ms_per_step"], r["trace_ms_per_step"], r["shade_ms_per_step"], r["frac"], r["trace_launches_per_step"], l["e2e"]["value"]))
EOF
done
