#!/bin/bash
# Round 2, final GPU-box visit: all -m gpu tests, both bench arms of the default line (config 5), the ncu launch list of the
# bench command, DRAM traffic of every k_trace launch of the timed step, --set full captures of mid-step k_trace<3,1,0> and
# k_shade launches, config 2 bench line.  Usage (under gpurun): bash tools/gpu_r2final.sh <tag>
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_$TAG.err; cut -c1-300 $OUT/bench_ref_$TAG.json
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg2_$TAG.json 2>> $OUT/bench_$TAG.err; cut -c1-300 $OUT/bench_cfg2_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
KT='regex:k_trace<.int.3, .bool.1, .bool.0>'
$CMD > $OUT/plain_$TAG.log 2>&1 || exit 1
NTL=$(python -c "import json;print(int(json.loads(open('$OUT/plain_$TAG.log').read().strip().splitlines()[-1])['roofline']['trace_launches_per_step']))" 2>/dev/null || echo 32)
echo "trace launches per step $NTL"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
# the counting pass + 3 warm-up steps + 2 e2e warm-ups come before / after: the timed step's launches are [4 NTL, 5 NTL)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --kernel-name-base demangled -k "$KT" -s $((4 * NTL)) -c $NTL --csv \
    --log-file $OUT/trace_traffic_$TAG.csv $CMD > $OUT/ncu_traffic_$TAG.log 2>&1
echo "ncu traffic rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$KT" -s $((4 * NTL + NTL / 3)) -c 2 -f -o $OUT/prof_trace_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_shade' -s $((5 * (4 * NTL + NTL / 3))) -c 5 -f -o $OUT/prof_shade_$TAG $CMD > $OUT/ncu_fulls_$TAG.log 2>&1
echo "ncu full shade rc=$?"
ls -la $OUT/*$TAG* | tail -14
