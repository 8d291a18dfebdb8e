#!/bin/bash
# bench.py at the shard sizes of a 1/4/8-GPU split on one GPU (device-resident leg only)
mkdir -p gpurun_out
for n in ${SIZES:-100000000 25000000 12500000}; do
  python bench.py --n $n --steps 40 --no-cpu-baseline --e2e-fits 0 > gpurun_out/bench_p_$n.json 2> gpurun_out/bench_p_$n.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_p_$n.json'))
r=d['roofline']
print('n %d value %.1f ms/step %.4f pass_ms %.4f stream_us %.1f step_us %.1f host %s launches %d' % (d['config']['n_per_gpu'], d['value'], d['ms_per_step'], r['avg_launch_ms'], r['stream_us'], r['step_us'], {k: round(v) for k, v in d['config']['host_us_per_fit'].items()}, d['gpu_launches']))
"
done
