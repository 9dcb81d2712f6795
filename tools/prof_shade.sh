#!/bin/bash
# launch list of one 3-spp render + ncu --set full of the first-bounce shade kernels (second render: skip the first render's 45 shade launches)
TAG=${1:-r1}
python tools/prof_one.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_one_$TAG.csv python tools/prof_one.py > gpurun_out/prof_ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/prof_one.py > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shade -s ${2:-45} -c ${3:-10} -f -o gpurun_out/prof_shade_$TAG python tools/prof_one.py > gpurun_out/prof_ncu_s_$TAG.log 2>&1
echo "shade rc=$?"; cat gpurun_out/prof_plain_$TAG.log
