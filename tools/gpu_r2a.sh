#!/bin/bash
# Round 2, GPU visit A: all -m gpu tests (new parity tests included), the new default bench line (config 5) + config 2,
# A/B of the round-1 library vs the cull / streams changes on the instanced scene and on config 2, and ncu --set full of
# mid-render k_trace<INST> / k_shade launches of the instanced scene.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -25 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_cfg5_$TAG.json 2> $OUT/bench_cfg5_$TAG.err; echo "bench cfg5 rc=$?"; cat $OUT/bench_cfg5_$TAG.json; tail -3 $OUT/bench_cfg5_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_cfg5_$TAG.json 2>> $OUT/bench_cfg5_$TAG.err; cat $OUT/bench_ref_cfg5_$TAG.json
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg2_$TAG.json 2> $OUT/bench_cfg2_$TAG.err; echo "bench cfg2 rc=$?"; cat $OUT/bench_cfg2_$TAG.json
CFG=3 STEP=8 bash tools/ab2.sh r1 nocull nostreams cur 2>&1 | tee $OUT/ab_cfg3_$TAG.txt
CFG=2 STEP=16 bash tools/ab2.sh r1 nostreams cur 2>&1 | tee $OUT/ab_cfg2_$TAG.txt
CMD="python bench.py --config 3 --step-spp 8 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_cfg3_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_trace<.int.3, .bool.1, .bool.0>' -s 60 -c 2 -f -o $OUT/prof_trace_cfg3_$TAG $CMD > $OUT/ncu_trace_cfg3_$TAG.log 2>&1
echo "ncu trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_shade' -s 250 -c 5 -f -o $OUT/prof_shade_cfg3_$TAG $CMD > $OUT/ncu_shade_cfg3_$TAG.log 2>&1
echo "ncu shade rc=$?"
ls -la $OUT | tail -15
