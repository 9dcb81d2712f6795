#!/bin/bash
# Round 2, GPU visit J: ambientocclusion integrator parity + throughput, then the whole GPU suite
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ao_integrator or aov" 2>&1 | tail -5
python tools/bench_integrators.py --only ao > $OUT/integrators_r2j.jsonl 2> $OUT/integrators_r2j.err; cat $OUT/integrators_r2j.jsonl; tail -3 $OUT/integrators_r2j.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu_r2j.txt
