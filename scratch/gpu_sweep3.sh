#!/bin/bash
mkdir -p gpurun_out
export GSLNLS_TRACE_E2E=1
NS=1e8,1.25e7 KEEPS=0 TUNES="tiled=0,block=256,unroll=4,minb=2;prefetch=1,unroll=2,block=256,minb=2;prefetch=1,unroll=2,block=192,minb=3;prefetch=1,unroll=2,block=288,minb=2;prefetch=1,unroll=2,block=128,minb=4;prefetch=1,unroll=3,block=256,minb=2;prefetch=1,unroll=1,block=256,minb=3;prefetch=1,unroll=1,block=192,minb=4;prefetch=1,unroll=4,block=256,minb=2;prefetch=0,unroll=2,block=192,minb=3" timeout 900 python scratch/sweep_l2.py > gpurun_out/sweep_pf.log 2>&1
NS=2.5e7,1.25e7 KEEPS=0,32,64,96 TUNES="tiled=0,block=256,unroll=4,minb=2" timeout 600 python scratch/sweep_l2.py > gpurun_out/sweep_l2b.log 2>&1
