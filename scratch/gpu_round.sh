#!/bin/bash
# one gpurun call: GPU tests, bench, launch list, ncu --set full of K1 and K1b, e2e phase trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
ls MEASURED_PEAKS.json >> gpurun_out/gpu.txt 2>&1; cat MEASURED_PEAKS.json >> gpurun_out/gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
GSLNLS_TRACE_E2E=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-fits 3 > gpurun_out/bench_e2e_trace.json 2> gpurun_out/bench_e2e_trace.err
timeout 300 python scratch/exp_server.py > gpurun_out/exp_server.log 2>&1
timeout 300 python scratch/config4.py > gpurun_out/config4.log 2>&1
# ncu: kernels replay one at a time, so the resident server (kernels that wait on each other) is off
export GSLNLS_SERVER=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 24 --warmup 3 --no-cpu-baseline --e2e-fits 0 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nls_pass -s 4 -c 2 -o gpurun_out/k1_full -f \
  python bench.py --steps 24 --warmup 3 --no-cpu-baseline --e2e-fits 0 > gpurun_out/ncu_k1.log 2>&1
N=1e7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nls_pass -s 2 -c 1 -o gpurun_out/k1b_full -f \
  python scratch/config4.py > gpurun_out/ncu_k1b.log 2>&1
ls -la gpurun_out
