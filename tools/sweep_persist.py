#!/usr/bin/env python
"""does an L2 persisting set-aside make evict_last hints retain a shard's head across passes?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gslnls_b200 as G
theta = [4.0, 1.3, 0.9]
os.environ["GSLNLS_TUNE"] = "tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1"
for n in (6_250_000, 12_500_000, 25_000_000):
    x, y = bench.synth_rows(0, n, bench.N_FULL)
    for keep in (0, 40, 64, 80, 1000):
        os.environ["GSLNLS_L2_KEEP_MB"] = str(keep)
        m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
        pb = G.Problem(m, n).upload([x], y)
        pb.time_passes(theta, 20)
        ms = min(pb.time_passes(theta, 50) for _ in range(3))
        print("persist=%s n=%9d (%4.0f MB) keep %4d MB  pass %6.1f us  %6.0f GB/s" % (os.environ.get("GSLNLS_L2_PERSIST_MB", "-"), n, 16e-6 * n, keep, ms * 1e3, 16.0 * n / (ms * 1e-3) / 1e9), flush=True)
        pb.close()
