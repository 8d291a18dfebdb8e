#!/usr/bin/env python
"""Where does a pass kernel spend its time on a small shard?  Per-CTA phase stamps (gslnls_problem_trace) of one
launch, for the shard sizes of a 1/2/4/8-GPU split and both load paths.  Run through gpurun."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

VARIANTS = {"ldg": "tiled=0,block=256,unroll=3,minb=2,prefetch=1,fexp=1",
            "tma": "tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1"}


def main():
    import gslnls_b200 as G
    from gslnls_b200 import _lib
    L = _lib.lib()
    theta = [4.0, 1.3, 0.9]
    sizes = [int(s) for s in sys.argv[1:]] or [12_500_000, 25_000_000, 100_000_000]
    for n in sizes:
        x, y = bench.synth_rows(0, n, bench.N_FULL)
        for vname, tune in VARIANTS.items():
            os.environ["GSLNLS_TUNE"] = tune
            m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
            pb = G.Problem(m, n).upload([x], y)
            pb.eval_packet(theta)
            ms = min(pb.time_passes(theta, 50) for _ in range(3))
            L.gslnls_problem_trace(pb.handle, 1, None, 0, None)
            for _ in range(3):
                pb.time_passes(theta, 4)   # back-to-back launches; the stamps of the last one remain
            buf = (C.c_uint64 * (32 * 1024))()
            nc = C.c_int()
            L.gslnls_problem_trace(pb.handle, 1, buf, 1024, C.byref(nc))
            t = np.array(buf[:32 * nc.value], dtype=np.float64).reshape(nc.value, 32)
            t0 = t[:, 0].min()
            rel = (t - t0) * 1e-3
            last = int(np.argmax(t[:, 5]))
            q = lambda a: "min %6.1f med %6.1f max %6.1f" % (a.min(), np.median(a), a.max())  # noqa: E731
            print("n=%9d %-3s ctas %3d  pass %6.1f us (back to back)" % (n, vname, nc.value, ms * 1e3))
            print("   entry            %s" % q(rel[:, 0]))
            print("   request seen     %s" % q(rel[:, 1]))
            print("   thread0 streamed %s" % q(rel[:, 2]))
            print("   cta streamed     %s" % q(rel[:, 3]))
            print("   partial written  %s" % q(rel[:, 4]))
            print("   packet published %6.1f (cta %d)   stream span per cta: %s" % (
                rel[last, 5], last, q(rel[:, 3] - rel[:, 1])), flush=True)
            nw = int(tune.split("block=")[1].split(",")[0]) // 32
            w = rel[:, 8:8 + nw]
            print("   warps done       " + " ".join("%5.1f" % v for v in np.median(w, axis=0)) + "   (median over ctas, per warp)")
            print("   slowest warp - fastest warp per cta: %s" % q(w.max(axis=1) - w.min(axis=1)), flush=True)
            L.gslnls_problem_trace(pb.handle, 0, None, 0, None)
            pb.close()


if __name__ == "__main__":
    main()
