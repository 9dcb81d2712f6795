#!/bin/bash
# Round 2, last GPU-box visit (after the cameras, map lights, resampled environment maps and the random sampler's sample
# arrays): all -m gpu tests, both bench arms of the default line (config 5), config 2, the ncu launch list of the bench command
TAG=${1:-r2fin2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.txt
tail -4 $OUT/${TAG}_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench_reference.json
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cfg2.json 2>> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench_cfg2.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/${TAG}_launches_bench.csv $CMD > /tmp/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
