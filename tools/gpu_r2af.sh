#!/bin/bash
# Round 2, GPU visit AF (2 GPUs): the final tree's multi-rank path - the default bench line at N = 2 (fixed job split over the
# ranks, NCCL film reduce) and the 2-rank command-line test
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2af_bench_n2.json 2> $OUT/r2af_bench_n2.err; echo "n2 rc=$?"
tail -1 $OUT/r2af_bench_n2.json | cut -c1-300
timeout 600 python -m pytest tests -m gpu -x -q -k "two_rank or gpus or cli" > $OUT/r2af_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2af_pytest.txt
