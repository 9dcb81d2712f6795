#!/bin/bash
# Round 2, GPU visit AE: A/B of the whole tree before Material "mix" / Sampler "stratified" (ab_old/ = an export of commit
# bf417d8 with its own library) against the current tree, alternating, on the two bench configurations; then the new
# sampler's GPU tests
OUT=gpurun_out
mkdir -p $OUT
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['rooflines']
print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms  trace', round(r['k_trace']['ms_per_step'],2), 'shade', round(r['k_shade']['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"; }
for rep in 1 2; do
  for cfg in 5 2; do
    (cd ab_old && python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | line "old cfg$cfg") | tee -a $OUT/r2ae_ab.txt
    python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | line "new cfg$cfg" | tee -a $OUT/r2ae_ab.txt
  done
done
timeout 900 python -m pytest tests -m gpu -x -q -k "stratified or random or ao_integrator or tree_integrators" > $OUT/r2ae_pytest.txt 2>&1
echo "pytest rc=$?"; tail -3 $OUT/r2ae_pytest.txt
