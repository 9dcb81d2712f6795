#!/bin/bash
# Round 2, GPU visit D: cooperative instance tests (parity on every instanced test, then A/B)
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py tests/test_oracle_pins.py -m gpu -q -k "instances or config3 or config5 or aov or exhaustive or lbvh" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -30 $OUT/pytest_gpu_$TAG.log
CFG=3 STEP=8 timeout 600 bash tools/ab2.sh nocoop coop coop_mb6 2>&1 | tee $OUT/ab_cfg3_$TAG.txt
