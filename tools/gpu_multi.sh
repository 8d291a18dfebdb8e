#!/bin/bash
# multi-GPU validation: tests + bench at N = $1 (run with gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
export GSLNLS_WATCHDOG_S=20
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$N.log 2>&1; tail -5 gpurun_out/pytest_multi_$N.log
for g in $(seq 1 $N); do
  if [ $g -eq 1 ] || [ $g -eq 2 ] || [ $g -eq 4 ] || [ $g -eq 8 ]; then
    if [ $g -eq 1 ]; then
      timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$g.json 2> gpurun_out/bench_n$g.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 40 --warmup 3 > gpurun_out/bench_n$g.json 2> gpurun_out/bench_n$g.err
    fi
    python -c "
import json
d=json.loads(open('gpurun_out/bench_n$g.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N=%d value %.1f ms/step %.4f pass_ms %.4f stream_us %.1f step_us %.1f e2e %.1f (%.1f ms/fit) host %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['avg_launch_ms'], r['stream_us'], r['step_us'], d['e2e']['value'], d['e2e']['ms_per_fit'], {k: round(v) for k, v in d['config']['host_us_per_fit'].items()}))
" || tail -5 gpurun_out/bench_n$g.err
  fi
done
