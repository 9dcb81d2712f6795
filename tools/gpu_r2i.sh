#!/bin/bash
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
CFG=2 STEP=16 timeout 300 bash tools/ab2.sh regs halton8 2>&1 | tee -a $OUT/ab_$TAG.txt
CFG=3 STEP=8 timeout 300 bash tools/ab2.sh regs halton8 2>&1 | tee -a $OUT/ab_$TAG.txt
CFG=1 STEP=0 STEPS=5 timeout 300 bash tools/ab2.sh regs halton8 2>&1 | tee -a $OUT/ab_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_oracle_pins.py -m gpu -q -k "sampler or golden or elementary or li_matches" 2>&1 | tail -3
