#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench_aux.py --workload gaussmix48 > gpurun_out/bench_aux_gaussmix48.json 2> gpurun_out/bench_aux.err
timeout 300 python bench_aux.py --workload mstart8192 > gpurun_out/bench_aux_mstart8192.json 2>> gpurun_out/bench_aux.err
timeout 300 python bench.py --algorithm lmaccel --no-cpu-baseline --e2e-fits 1 > gpurun_out/bench_lmaccel.json 2>> gpurun_out/bench_final.err
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
