#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
for cfg in "4 1" "1 1" "4 0" "1 0" "1000 1" "1000 0"; do
  set -- $cfg
  echo "== PROF_STRIDE=$1 PDL=$2" >> gpurun_out/bench5.log
  GSLNLS_PROF_STRIDE=$1 GSLNLS_PDL=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-fits 2 >> gpurun_out/bench5.log 2>> gpurun_out/bench5.err
done
for n in 25000000 12500000; do
for cfg in "4 1" "4 0" "1000 1" "1000 0"; do
  set -- $cfg
  echo "== n=$n PROF_STRIDE=$1 PDL=$2" >> gpurun_out/bench5.log
  GSLNLS_PROF_STRIDE=$1 GSLNLS_PDL=$2 timeout 300 python bench.py --n $n --steps 40 --warmup 5 --no-cpu-baseline --e2e-fits 2 >> gpurun_out/bench5.log 2>> gpurun_out/bench5.err
done
done
