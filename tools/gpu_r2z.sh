#!/bin/bash
# Round 2, GPU visit Z: ncu --set full of k_ri (whitted and directlighting "all" on the 1 M-triangle scene), five consecutive
# mid-step launches each; the reports are summarised on the box (they exceed what gpurun copies back)
OUT=gpurun_out
mkdir -p $OUT
for c in whitted direct_all; do
  CMD="python tools/bench_integrators.py --no-cpu --steps 1 --only $c"
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_ri" -s 30 -c 5 -f -o /tmp/prof_${c}_r2z $CMD > $OUT/ncu_${c}_r2z.log 2>&1
  echo "$c ncu rc=$?"
  python tools/ncu_summary.py /tmp/prof_${c}_r2z.ncu-rep | cut -c1-260 > $OUT/r2z_ncu_${c}_summary.txt
done
