#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
rm -f gpurun_out/bench6.log
for n in 100000000 25000000 12500000; do
  echo "== n=$n" >> gpurun_out/bench6.log
  timeout 300 python bench.py --n $n --steps 40 --warmup 5 --no-cpu-baseline --e2e-fits 2 >> gpurun_out/bench6.log 2>> gpurun_out/bench6.err
done
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
