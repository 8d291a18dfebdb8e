#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/bench7.log
for rep in 1 2; do
for n in 12500000 25000000; do
for e in 0 1; do
  echo "== n=$n NO_EARLY=$e rep=$rep" >> gpurun_out/bench7.log
  GSLNLS_NVRTC_FLAGS="-DNLS_NO_EARLY=$e" timeout 300 python bench.py --n $n --steps 80 --warmup 5 --no-cpu-baseline --e2e-fits 0 >> gpurun_out/bench7.log 2>> gpurun_out/bench7.err
done
done
done
