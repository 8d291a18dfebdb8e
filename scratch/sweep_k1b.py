"""scratch: does K1 time depend on warm-up duration / theta / allocation?"""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gslnls_b200 import Model, Problem
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
torch.cuda.set_device(0)
model = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=False)
n = 100_000_000
x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
pb = Problem(model, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
for th in ([4.0, 1.3, 0.9], [1.0, 1.0, 0.0], [5.0, 1.5, 1.0]):
    th = np.array(th)
    for rep in range(6):
        ms = pb.time_passes(th, 200)
        print("theta", th, "rep", rep, "pass %.1f us" % (ms * 1e3), clocks(), flush=True)
# library-owned buffers uploaded from host
xh, yh = x.cpu().numpy(), y.cpu().numpy()
pb2 = Problem(model, n, False, 0).upload([xh], yh)
for rep in range(3):
    print("uploaded: pass %.1f us" % (1e3 * pb2.time_passes(np.array([4.0, 1.3, 0.9]), 200)), clocks(), flush=True)
