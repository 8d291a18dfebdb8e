#!/bin/bash
# multi-GPU check: N = number of GPUs of this box to use
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/pytest_multi_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
for k in 1 $N; do
  if [ $k -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $k --steps 40 --warmup 5 > gpurun_out/bench_g$k.json 2> gpurun_out/bench_g$k.err
  fi
done
