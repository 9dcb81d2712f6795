#!/bin/bash
# Round 2, GPU visit M (8 GPUs): fixed-job strong scaling of the default bench line (config 5, 8 sample numbers per step) at N = 1, 2, 4, 8
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi_r2m.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_n1_r2m.json 2> $OUT/bench_n1_r2m.err; echo "n1 rc=$?"
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_n${N}_r2m.json 2> $OUT/bench_n${N}_r2m.err; echo "n$N rc=$?"
done
for N in 1 2 4 8; do tail -1 $OUT/bench_cfg5_n${N}_r2m.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']
print('N=$N', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"; done | tee $OUT/scale_r2m.txt
