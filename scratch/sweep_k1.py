"""scratch: K1 pass time vs n and tune (launch-ordered back-to-back passes)"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("PYTHONPATH", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))).split(":")[0])
import torch
from gslnls_b200 import Model, Problem

torch.cuda.set_device(0)
rhs = os.environ.get("RHS", "A * exp(-lam * x) + b")
names = os.environ.get("PARS", "A,lam,b").split(",")
theta = np.array([float(v) for v in os.environ.get("THETA", "4.0,1.3,0.9").split(",")])
model = Model(rhs, names, ["x"], jac=True, fvv=False)
for n in [int(float(v)) for v in os.environ.get("NS", "1.25e7,2.5e7,5e7,1e8").split(",")]:
    x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
    y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
    pb = Problem(model, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
    pb.time_passes(theta, 5)
    ms = min(pb.time_passes(theta, 20) for _ in range(3))
    print("n=%d  pass %.1f us  %.0f GB/s (ideal %.1f us at 6451 GB/s)" % (n, ms * 1e3, 16.0 * n / ms / 1e6, 16.0 * n / 6451.2e3), flush=True)
    pb.close()
    del x, y
