#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "8" ) > gpurun_out/pytest_multi_8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_8.log
for k in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k bench.py --gpus $k --steps 40 --warmup 5 > gpurun_out/bench_h$k.json 2> gpurun_out/bench_h$k.err
done
