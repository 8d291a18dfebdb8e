#!/usr/bin/env python
"""K1b (tiled DMMA pass, p = 48): who waits for whom?  Cycles every warp spends waiting on its mbarrier
(consumers on `full`, producers on `empty`) against the cycles of its streaming loop.  Run through gpurun."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import gslnls_b200 as G
    from gslnls_b200 import _lib
    L = _lib.lib()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    x, y = bench.gaussmix_rows(0, n, n)
    rhs, names = bench.gaussmix_formula(16)
    start = bench.gaussmix_start(16)
    m = G.Model(rhs, names, ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], y)
    pb.eval_packet(start)
    ms = min(pb.time_passes(start, 10) for _ in range(3))
    L.gslnls_problem_trace(pb.handle, 1, None, 0, None)
    pb.time_passes(start, 3)
    buf = (C.c_uint64 * (32 * 1024))()
    nc = C.c_int()
    L.gslnls_problem_trace(pb.handle, 1, buf, 1024, C.byref(nc))
    t = np.array(buf[:32 * nc.value], dtype=np.float64).reshape(nc.value, 32)
    wait, total = t[:, :16], t[:, 16:32]
    frac = wait / np.maximum(total, 1)
    print("n=%d p=48 pass %.1f us, %d CTAs; loop cycles median %.0f" % (n, ms * 1e3, nc.value, np.median(total)))
    print("consumer warps 0-3   wait fraction: " + " ".join("%.3f" % v for v in np.median(frac[:, :4], axis=0)))
    print("producer warps 4-15  wait fraction: " + " ".join("%.3f" % v for v in np.median(frac[:, 4:16], axis=0)))
    flops = n * (48 * 49 + 98.0)
    print("algorithmic %.1f TFLOP/s" % (flops / (ms * 1e-3) / 1e12))
    pb.close()


if __name__ == "__main__":
    main()
