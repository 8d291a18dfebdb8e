#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 900 python scratch/sweep_model2.py > gpurun_out/sweep_model2.log 2>&1
