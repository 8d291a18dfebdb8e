#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "$N" ) > gpurun_out/pytest_multi_q$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_q$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_q$N.json 2> gpurun_out/bench_q$N.err
