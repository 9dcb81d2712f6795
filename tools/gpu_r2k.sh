#!/bin/bash
# Round 2, GPU visit K: directlighting / whitted parity + throughput of the integrators beside "path"
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tree_integrators or ao_integrator" 2>&1 | tail -15
python tools/bench_integrators.py > $OUT/integrators_r2k.jsonl 2> $OUT/integrators_r2k.err; cat $OUT/integrators_r2k.jsonl; tail -3 $OUT/integrators_r2k.err
