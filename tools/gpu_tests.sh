#!/bin/bash
# GPU test-suite without -x (all failures in one go) + optional extra command
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -40
if [ -n "$1" ]; then bash -c "$1"; fi
