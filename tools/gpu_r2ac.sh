#!/bin/bash
# Round 2, GPU visit AC: orthographic / environment cameras - parity test, then the default bench line (k_raygen gained a camera switch)
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "cameras or camera" > $OUT/r2ac_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 $OUT/r2ac_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r2ac_bench.json 2> $OUT/r2ac_bench.err
echo "bench rc=$?"; cut -c1-400 $OUT/r2ac_bench.json
