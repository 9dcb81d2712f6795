#!/bin/bash
# Round 2, GPU visit G: ncu --set full of the steady-state shade kernels on config 2 (per-line view for plastic / metal / glass)
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --config 2 --step-spp 16 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_cfg2_$TAG.log 2>&1 && tail -c 600 $OUT/plain_cfg2_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_shade' -s 260 -c 5 -f -o $OUT/prof_shade_cfg2_$TAG $CMD > $OUT/ncu_shade_cfg2_$TAG.log 2>&1
echo "ncu shade rc=$?"
ls -la $OUT/*.ncu-rep | tail -3
