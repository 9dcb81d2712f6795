#!/bin/bash
# Round 2, GPU visit AD: Material "mix" (scaled lobes in the five-lobe class) - parity tests of the materials, then the two
# bench lines again (dbsdf.cuh changed: the BASELINE classes must not have moved)
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "mix or extra_materials or tree_integrators or baseline" > $OUT/r2ad_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 $OUT/r2ad_pytest.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2ad_bench.json 2> $OUT/r2ad_bench.err; cut -c1-200 $OUT/r2ad_bench.json
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2ad_bench_cfg2.json 2>> $OUT/r2ad_bench.err; cut -c1-200 $OUT/r2ad_bench_cfg2.json
