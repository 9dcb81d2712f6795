#!/bin/bash
# Round 2, GPU visit H: shade kernels with one aggregated append per vertex and a register-resident BsdfD: parity, then A/B
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_branches.py tests/test_baseline_configs.py tests/test_golden.py -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -8 $OUT/pytest_gpu_$TAG.log
CFG=2 STEP=16 timeout 300 bash tools/ab2.sh async regs 2>&1 | tee -a $OUT/ab_$TAG.txt
CFG=3 STEP=8 timeout 300 bash tools/ab2.sh async regs 2>&1 | tee -a $OUT/ab_$TAG.txt
CFG=1 STEP=0 STEPS=5 timeout 300 bash tools/ab2.sh async regs 2>&1 | tee -a $OUT/ab_$TAG.txt
