#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scratch/sweep_tma.py > gpurun_out/sweep_tma.log 2>&1
GSLNLS_TUNE="tiled=2,block=288,unroll=2,minb=2,stages=4" timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_tma.log 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
GSLNLS_TRACE_E2E=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-fits 4 > gpurun_out/bench2_trace.json 2> gpurun_out/bench2_trace.err
