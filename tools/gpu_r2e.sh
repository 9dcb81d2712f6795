#!/bin/bash
# Round 2, GPU visit E (2 GPUs): fixed-job strong scaling N=1 vs N=2 (config 5 and config 2), NCCL 2-rank CLI test
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi_$TAG.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_n1_$TAG.json 2> $OUT/bench_n1_$TAG.err; echo "n1 rc=$?"; cat $OUT/bench_cfg5_n1_$TAG.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_cfg5_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err; echo "n2 rc=$?"; cat $OUT/bench_cfg5_n2_$TAG.json | cut -c1-400; tail -3 $OUT/bench_n2_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 2 --steps 3 --warmup 3 > $OUT/bench_cfg2_n2_$TAG.json 2>> $OUT/bench_n2_$TAG.err; echo "cfg2 n2 rc=$?"; cat $OUT/bench_cfg2_n2_$TAG.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_n2_$TAG.json 2>> $OUT/bench_n2_$TAG.err; echo "ref n2 rc=$?"; cat $OUT/bench_ref_n2_$TAG.json | cut -c1-200
timeout 600 python -m pytest tests/test_cli.py -m gpu -q 2>&1 | tail -3
