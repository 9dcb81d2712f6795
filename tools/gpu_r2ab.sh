#!/bin/bash
# Round 2, GPU visit AB: random-sampler sample arrays (ambientocclusion, directlighting "all") - parity tests, then the
# integrator throughput table again (k_ri / k_ao gained a sampler branch)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -k "ao_integrator or tree_integrators or random" > $OUT/r2ab_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 $OUT/r2ab_pytest.txt
timeout 900 python tools/bench_integrators.py --no-cpu > $OUT/r2ab_integrators.jsonl 2> $OUT/r2ab_integrators.err
echo "bench rc=$?"; cat $OUT/r2ab_integrators.jsonl | cut -c1-300
