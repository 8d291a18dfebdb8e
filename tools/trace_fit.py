#!/usr/bin/env python
"""Per-CTA phase stamps of ONE pass inside a running fit (resident server + persistent pass kernel): where the
microseconds of a small-shard step go.  Run through gpurun."""
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
    sizes = [int(s) for s in sys.argv[1:]] or [12_500_000, 100_000_000]
    for n in sizes:
        x, y = bench.synth_rows(0, n, bench.N_FULL)
        m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
        pb = G.Problem(m, n).upload([x], y)
        pb.fit(list(bench.START))
        L.gslnls_problem_trace(pb.handle, 1, None, 0, None)
        for npass in (5, 6):
            pb.fit_begin(list(bench.START))
            pb.fit_run(npass)
            buf = (C.c_uint64 * (32 * 1024))()
            nc = C.c_int()
            L.gslnls_problem_trace(pb.handle, 1, buf, 1024, C.byref(nc))
            st, sp, cnt = pb.channel_stats()
            pb.fit_end()
            t = np.array(buf[:32 * nc.value], dtype=np.float64).reshape(nc.value, 32)
            t0 = t[:, 1].min()
            rel = (t - t0) * 1e-3
            last = int(np.argmax(t[:, 5]))
            q = lambda a: "min %6.2f med %6.2f max %6.2f" % (a.min(), np.median(a), a.max())  # noqa: E731
            nw = 12
            w = rel[:, 8:8 + nw]
            print("n=%9d ctas %3d  pass #%d of a fit   (us after the first CTA saw the request)" % (n, nc.value, npass))
            print("   request seen     %s" % q(rel[:, 1]))
            print("   warps streamed   %s   per-cta spread %s" % (q(w.max(axis=1)), q(w.max(axis=1) - w.min(axis=1))))
            print("   cta streamed     %s" % q(rel[:, 3]))
            print("   partial written  %s" % q(rel[:, 4]))
            print("   packet published %6.2f (cta %d)" % (rel[last, 5], last), flush=True)
        L.gslnls_problem_trace(pb.handle, 0, None, 0, None)
        pb.close()


if __name__ == "__main__":
    main()
