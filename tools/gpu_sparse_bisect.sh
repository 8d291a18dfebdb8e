#!/bin/bash
T="tests/test_gpu_sparse.py::test_grouped_exponential_vs_oracle"
run() { echo "== $*"; env "$@" timeout 100 python -m pytest $T -q -x --timeout 60 2>&1 | grep -E "passed|failed|^E  " | head -4; }
for v in d e; do
  cd scratch/v$v
  run V=$v; run V=$v; run V=$v GSLNLS_SP_MINB=2
  cd ../..
done
