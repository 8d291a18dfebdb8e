"""scratch: config 4 (sum of 16 Gaussians, p=48, n=1e7) pass time and fit"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gaussmix import gaussmix_formula, gaussmix_truth
from gslnls_b200 import Model, Problem
K = 16; p = 48
n = int(float(os.environ.get("N", "1e7")))
torch.cuda.set_device(0)
x = torch.linspace(0, 100, n, dtype=torch.float64, device="cuda")
th = gaussmix_truth(K)
y = torch.zeros_like(x)
for k in range(K):
    y += th[3*k] * torch.exp(-((x - th[3*k+1]) ** 2) / th[3*k+2] ** 2)
y += 0.5 * torch.randn(n, dtype=torch.float64, device="cuda")
rhs, names = gaussmix_formula(K)
m = Model(rhs, names, ["x"], jac=True)
pb = Problem(m, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
start = th * (1.0 + 0.02 * (-1.0) ** np.arange(p))
pb.time_passes(start, 2)
ms = min(pb.time_passes(start, 5) for _ in range(3))
flops = n * (p * (p + 1) + 2 * p + 2.0)
print("n=%d pass %.3f ms  SYRK-algorithmic %.2f TFLOP/s (FP64 peak 37.1)" % (n, ms, flops / ms / 1e9), flush=True)
for alg in ("dogleg", "ddogleg", "lm"):
    t0 = time.perf_counter()
    r = pb.fit(start, algorithm=alg)
    dt = time.perf_counter() - t0
    print(alg, r["status"], "niter", r["niter"], "npass", r["npass"], "ssr %.6g" % r["ssr"], "wall %.1f ms" % (dt * 1e3),
          "max rel par err vs truth %.2e" % np.max(np.abs(r["par"] / th - 1)), flush=True)
