#!/bin/bash
# ncu full captures of the two dominant kernels on the bench workload at 8 spp (same kernels, fewer waves).
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace -s ${2:-27} -c 2 -f -o $OUT/prof_trace_$TAG $CMD > $OUT/ncu_trace_$TAG.log 2>&1
echo "ncu trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade -s ${3:-27} -c 2 -f -o $OUT/prof_shade_$TAG $CMD > $OUT/ncu_shade_$TAG.log 2>&1
echo "ncu shade rc=$?"
cat $OUT/plain_$TAG.log | tail -2
