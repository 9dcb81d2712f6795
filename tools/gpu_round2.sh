#!/bin/bash
# One GPU-box visit (pipelined bounce loop): parity tests, both bench arms, the ncu launch list of the bench command, the
# DRAM traffic of every k_trace launch of the timed step, and one --set full capture of mid-step k_trace / k_shade launches.
# Usage (under gpurun): bash tools/gpu_round2.sh <tag> [skip_tests]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
  tail -5 $OUT/pytest_gpu_$TAG.log
fi
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_ref_$TAG.json
python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
KT='regex:k_trace<.int.3, .bool.0, .bool.0>'
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
NTL=$(python -c "import json;print(int(json.load(open('$OUT/plain_$TAG.log'))['roofline']['trace_launches_per_step']))" 2>/dev/null || echo 42)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --kernel-name-base demangled -k "$KT" -s $((3 * NTL)) -c $NTL --csv \
    --log-file $OUT/trace_traffic_$TAG.csv $CMD > $OUT/ncu_traffic_$TAG.log 2>&1
echo "ncu traffic rc=$? (launches per step $NTL)"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$KT" -s $((3 * NTL + NTL / 2)) -c 3 -f -o $OUT/prof_trace_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full trace rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_shade' -s $((5 * (3 * NTL + NTL / 2))) -c 5 -f -o $OUT/prof_shade_$TAG $CMD > $OUT/ncu_fulls_$TAG.log 2>&1
echo "ncu full shade rc=$?"
ls -la $OUT | tail -12
