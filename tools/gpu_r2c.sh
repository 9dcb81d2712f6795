#!/bin/bash
# Round 2, GPU visit C: new -m gpu tests (2-rank CLI, oracle pins), k_trace variants on the instanced scene
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cli.py tests/test_oracle_pins.py tests/test_branches.py tests/test_baseline_configs.py -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -30 $OUT/pytest_gpu_$TAG.log
CFG=3 STEP=8 bash tools/ab2.sh inl1 noflat flat flat_mb6 2>&1 | tee $OUT/ab_cfg3_$TAG.txt
CFG=2 STEP=16 bash tools/ab2.sh inl1 flat 2>&1 | tee $OUT/ab_cfg2_$TAG.txt
