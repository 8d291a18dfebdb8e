#!/bin/bash
# one GPU-box round of the usual checks; everything lands in gpurun_out/ (run through gpurun)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"
tail -3 gpurun_out/ncu_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
timeout 600 python bench.py --config gaussmix48 > gpurun_out/bench_gaussmix48.json 2> gpurun_out/bench_gaussmix48.err; echo "bench g48 rc=$?"; cat gpurun_out/bench_gaussmix48.json
timeout 600 python bench.py --config mstart8192 > gpurun_out/bench_mstart8192.json 2> gpurun_out/bench_mstart8192.err; echo "bench ms rc=$?"; cat gpurun_out/bench_mstart8192.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
