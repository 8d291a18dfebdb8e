#!/bin/bash
# A/B of the solver kernel's residency (GSLNLS_SP_MINB) on the sparse bench workload, after the parity tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sparse.py -q -x --timeout 90 2>&1 | tail -4 | cut -c1-300
timeout 100 python -m pytest tests/test_gpu_sparse.py -q -x --timeout 90 -k "grouped" 2>&1 | tail -2 | cut -c1-300
for m in 4 3 2; do
  echo "== GSLNLS_SP_MINB=$m"
  GSLNLS_SP_MINB=$m timeout 120 python bench.py --config sparse --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('ms/fit', d['ms_per_step'], d['config']['device_ms_per_fit'], 'frac', d['roofline']['frac'], d['config']['step'])
    else: print(ln.rstrip()[:300])
"
done
