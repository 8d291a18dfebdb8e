#!/bin/bash
# sparse path: parity tests, then the bench workload under the variants named on the command line (env assignments)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sparse.py -q -x --timeout 90 2>&1 | tail -4 | cut -c1-400
for m in "$@"; do
  echo "== $m"
  env $m timeout 120 python bench.py --config sparse --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('ms/fit', d['ms_per_step'], d['config']['device_ms_per_fit'], 'frac', d['roofline']['frac'], d['config']['step'][:90])
    else: print(ln.rstrip()[:300])
"
done
