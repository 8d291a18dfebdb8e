#!/bin/bash
# 8-GPU validation, kept short (charged 8x): world-8 tests, then bench at N = 8 and 4
mkdir -p gpurun_out
export GSLNLS_WATCHDOG_S=20
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8" > gpurun_out/pytest_multi_8.log 2>&1; tail -4 gpurun_out/pytest_multi_8.log
for g in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 40 --warmup 3 > gpurun_out/bench_n$g.json 2> gpurun_out/bench_n$g.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n$g.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N=%d value %.1f ms/step %.4f pass_ms %.4f stream_us %.1f step_us %.1f e2e %.1f (%.1f ms/fit) host %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['avg_launch_ms'], r['stream_us'], r['step_us'], d['e2e']['value'], d['e2e']['ms_per_fit'], {k: round(v) for k, v in d['config']['host_us_per_fit'].items()}))
" || tail -5 gpurun_out/bench_n$g.err
done
timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-fits 0 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N=1 value %.1f ms/step %.4f pass_ms %.4f stream_us %.1f step_us %.1f' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['stream_us'], r['step_us']))
"
