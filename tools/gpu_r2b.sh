#!/bin/bash
# Round 2, GPU visit B: A/B of k_trace (survivor loop) and k_shade variants (single-site inlining, L2-only global loads,
# multi-level CDF search, occupancy) on the instanced scene and on config 2.
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
CFG=3 STEP=8 bash tools/ab2.sh cur surv noinl1 inl1 noinl1cg inl1cg kary karycg kary_mb5 kary_mb6 2>&1 | tee $OUT/ab_cfg3_$TAG.txt
CFG=2 STEP=16 bash tools/ab2.sh cur inl1 inl1cg kary karycg kary_mb5 2>&1 | tee $OUT/ab_cfg2_$TAG.txt
