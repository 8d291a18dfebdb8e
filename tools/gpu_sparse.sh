#!/bin/bash
# sparse-row path: GPU tests, bench line, ncu launch list + one full capture of sp_step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sparse.py -q --timeout 120 2>&1 | grep -E "^E|passed|failed|^>" | head -20 | cut -c1-400
if [ "$1" != "nobench" ]; then
timeout 400 python bench.py --config sparse > gpurun_out/r02_bench_sparse.json 2> gpurun_out/r02_bench_sparse.err
echo "bench rc=$?"; cat gpurun_out/r02_bench_sparse.json; tail -5 gpurun_out/r02_bench_sparse.err
fi
if [ "$1" = "ncu" ] || [ "$2" = "ncu" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_sparse_launches.csv \
    python bench.py --config sparse --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_sparse_launches.err
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sp_step -s 2 -c 2 -o /tmp/r02_sparse_full -f \
    python bench.py --config sparse --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_sparse_full.err
echo "full rc=$?"
ncu -i /tmp/r02_sparse_full.ncu-rep --page raw --csv > gpurun_out/r02_sparse_full_raw.csv 2>/dev/null
ncu -i /tmp/r02_sparse_full.ncu-rep --page details --csv > gpurun_out/r02_sparse_full_details.csv 2>/dev/null
ncu -i /tmp/r02_sparse_full.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-12 > gpurun_out/r02_sparse_source.csv
ls -la gpurun_out/r02_sparse*
fi
