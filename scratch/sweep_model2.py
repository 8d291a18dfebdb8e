"""scratch: neighbourhood of the new K1 default (prefetch + table exp)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gslnls_b200 import Model, Problem
torch.cuda.set_device(0)
TUNES = [t for t in os.environ.get("TUNES", "").split(";")] if os.environ.get("TUNES") else [""]
FLAGS = os.environ.get("FLAGSETS", "").split(";")
th = np.array([4.0, 1.3, 0.9])
for n in [int(float(v)) for v in os.environ.get("NS", "1e8,1.25e7,1e6").split(",")]:
    x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
    y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
    for tune, flags in [(t, f) for t in TUNES for f in FLAGS]:
        os.environ["GSLNLS_TUNE"] = tune
        os.environ["GSLNLS_NVRTC_FLAGS"] = flags
        try:
            m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=False)
            pb = Problem(m, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
            pb.time_passes(th, 20)
            ms = min(pb.time_passes(th, 100) for _ in range(3))
            print("n=%d %-52s %-16s pass %.1f us  %.0f GB/s" % (n, tune or "(default)", flags, ms * 1e3, 16.0 * n / ms / 1e6), flush=True)
            pb.close()
        except Exception as e:  # noqa: BLE001
            print("n=%d %-52s FAILED %s" % (n, tune, e), flush=True)
    del x, y
