#!/usr/bin/env python
"""pageable-memory upload rate against the number of staging threads (GSLNLS_UPLOAD_THREADS)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gslnls_b200 as G
n = 100_000_000
x, y = bench.synth_rows(0, n, n)
m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
for th in (4, 8, 12, 16):
    os.environ["GSLNLS_UPLOAD_THREADS"] = str(th)
    pb = G.Problem(m, n)
    pb.upload([x], y); pb.eval_packet([1.0, 1.0, 0.0])
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        pb.upload([x], y)
        pb.eval_packet([1.0, 1.0, 0.0])      # synchronises the stream (adds one 0.25 ms pass)
        best = min(best, time.perf_counter() - t0)
    print("threads %2d: %.2f ms  %.1f GB/s" % (th, best * 1e3, 1.6 / best), flush=True)
    pb.close()
