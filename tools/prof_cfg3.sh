#!/bin/bash
# ncu --set full of mid-render k_trace launches of the instanced scene (configs[2]) at 8 spp
TAG=${1:-r1}
OUT=gpurun_out
CMD="python tools/bench_configs.py --only 3 --spp3 8 --steps 1"
$CMD > $OUT/cfg3_plain_$TAG.log 2>&1 && cat $OUT/cfg3_plain_$TAG.log &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_trace<.int.3, .bool.1, .bool.0>' -s ${2:-22} -c 2 -f -o $OUT/prof_trace_cfg3_$TAG $CMD > $OUT/ncu_cfg3_$TAG.log 2>&1
echo "ncu rc=$?"
