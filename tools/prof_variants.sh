#!/bin/bash
# ncu --set full capture of trace launch #1 of the second render (the first bounce: continuation + shadow + MIS rays) for each variant library
for v in "$@"; do
  export PBRT_R3_B200_LIB=$PWD/pbrt-r3_b200/variants/lib_$v.so
  python tools/prof_one.py > gpurun_out/prof_plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 10 -c 1 -f -o gpurun_out/prof_trace_$v python tools/prof_one.py > gpurun_out/prof_ncu_$v.log 2>&1
  echo "$v rc=$?"; cat gpurun_out/prof_plain_$v.log
done
