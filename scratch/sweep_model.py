"""scratch: what bounds K1 at p = 3 -- memory side or SM side?  Same loads, different arithmetic."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gslnls_b200 import Model, Problem
torch.cuda.set_device(0)
CASES = [
    # (label, formula, params, nvrtc flags, tune)
    ("exp3 library exp", "A * exp(-lam * x) + b", "", ""),
    ("exp3 table exp", "A * exp(-lam * x) + b", "-UNLS_FAST_EXP -DNLS_FAST_EXP=1", ""),
    ("exp3 poly exp", "A * exp(-lam * x) + b", "-UNLS_FAST_EXP -DNLS_FAST_EXP=2", ""),
    ("linear (no exp)", "A + lam * x + b * x * x", "", ""),
    ("exp3 library exp, block 512 x1", "A * exp(-lam * x) + b", "", "block=512,unroll=4,minb=1"),
    ("exp3 library exp, block 128 x4", "A * exp(-lam * x) + b", "", "block=128,unroll=4,minb=4"),
    ("exp3 table exp, pf u3", "A * exp(-lam * x) + b", "-UNLS_FAST_EXP -DNLS_FAST_EXP=1", "prefetch=1,unroll=3,block=256,minb=2"),
    ("exp3 table exp, TMA 416", "A * exp(-lam * x) + b", "-UNLS_FAST_EXP -DNLS_FAST_EXP=1", "tiled=2,block=416,unroll=2,minb=1,stages=6"),
    ("exp3 table exp, TMA 288x2", "A * exp(-lam * x) + b", "-UNLS_FAST_EXP -DNLS_FAST_EXP=1", "tiled=2,block=288,unroll=2,minb=2,stages=4"),
    ("linear, TMA 416", "A + lam * x + b * x * x", "", "tiled=2,block=416,unroll=2,minb=1,stages=6"),
    ("linear, TMA 288x2", "A + lam * x + b * x * x", "", "tiled=2,block=288,unroll=2,minb=2,stages=4"),
    ("linear, TMA 288x2 u4 s3", "A + lam * x + b * x * x", "", "tiled=2,block=288,unroll=4,minb=2,stages=3"),
]
th = np.array([4.0, 1.3, 0.9])
for n in [int(float(v)) for v in os.environ.get("NS", "1e8,1.25e7,1e6,1e4").split(",")]:
    x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
    y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
    for label, rhs, flags, tune in CASES:
        os.environ["GSLNLS_NVRTC_FLAGS"] = flags
        os.environ["GSLNLS_TUNE"] = tune
        try:
            m = Model(rhs, ["A", "lam", "b"], ["x"], jac=True, fvv=False)
            pb = Problem(m, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
            pb.time_passes(th, 20)
            ms = min(pb.time_passes(th, 100) for _ in range(3))
            print("n=%d %-36s pass %.1f us  %.0f GB/s" % (n, label, ms * 1e3, 16.0 * n / ms / 1e6), flush=True)
            pb.close()
        except Exception as e:  # noqa: BLE001
            print("n=%d %-36s FAILED %s" % (n, label, e), flush=True)
    del x, y
