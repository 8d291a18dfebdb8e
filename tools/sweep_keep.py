#!/usr/bin/env python
"""A/B sweep on a B200: pass time of K1 against the L2-resident head of the shard (GSLNLS_L2_KEEP_MB) for
shard sizes of a 1/2/4/8-GPU split of n = 1e8, both load paths.  Run through gpurun; prints one line per case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

VARIANTS = {"ldg": "tiled=0,block=256,unroll=3,minb=2,prefetch=1,fexp=1",
            "tma": "tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1"}


def main():
    import gslnls_b200 as G
    theta = [4.0, 1.3, 0.9]
    for n in (12_500_000, 25_000_000, 50_000_000, 100_000_000):
        x, y = bench.synth_rows(0, n, bench.N_FULL)
        for vname, tune in VARIANTS.items():
            os.environ["GSLNLS_TUNE"] = tune
            ref = None
            for keep in (0, 32, 48, 64, 80, 96, 112):
                os.environ["GSLNLS_L2_KEEP_MB"] = str(keep)
                m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
                pb = G.Problem(m, n).upload([x], y)
                pk = pb.eval_packet(theta)
                if ref is None:
                    ref = pk
                pb.time_passes(theta, 20)
                ms = min(pb.time_passes(theta, 50) for _ in range(3))
                print("n=%9d %-3s keep %3d MB  pass %7.1f us  %6.0f GB/s algorithmic  bitwise_same_as_keep0=%s" % (
                    n, vname, keep, ms * 1e3, 16.0 * n / (ms * 1e-3) / 1e9, bool(np.array_equal(pk, ref))), flush=True)
                pb.close()


if __name__ == "__main__":
    main()
