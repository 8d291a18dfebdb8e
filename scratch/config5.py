"""scratch: config 5 (BASELINE.json configs[4]) -- 8192 multi-start candidates of the two-exponential
mixture, n = 4096, p = 4, mstart_p = 5 LM iterations each + log det(J^T J) screen, one batch"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gslnls_b200 import Model, Problem
from oracle import oracle as O
from scipy.stats import qmc

n, S = 4096, int(os.environ.get("S", "8192"))
rng = np.random.Generator(np.random.Philox(key=5))
x = np.linspace(0, 10, n)
y = 3 * np.exp(-0.5 * x) + 2 * np.exp(-3 * x) + 0.05 * rng.standard_normal(n)
starts = qmc.Sobol(4, scramble=False).random(S + 1)[1:] * 10.0   # Sobol points in the [0,10]^4 box
m = Model("A1*exp(-l1*x)+A2*exp(-l2*x)", ["A1", "l1", "A2", "l2"], ["x"], jac=True)
pb = Problem(m, n).upload([x], y)
out = pb.fit_batch(starts, iters=5)
ts = []
for rep in range(5):
    t0 = time.perf_counter()
    out = pb.fit_batch(starts, iters=5)
    ts.append(time.perf_counter() - t0)
best = int(np.nanargmin(np.where(np.isfinite(out["ssr"]), out["ssr"], np.inf)))
print("S=%d n=%d: fit_batch wall %.2f ms (min of 5; includes start upload + result download), %.0f candidate-iterations/s"
      % (S, n, min(ts) * 1e3, S * 5 / min(ts)), flush=True)
print("best candidate", best, "par", out["par"][best], "ssr %.6g" % out["ssr"][best], flush=True)
# CPU oracle on a sample of candidates, single thread
k = 64
t0 = time.perf_counter()
bad = 0
for c in range(k):
    ref = O.nls_large("expmix2", y, starts[c], x=x, algorithm="lm", maxiter=5)
    if ref["conv"] in (0, 11):
        if not (np.allclose(out["par"][c], ref["par"], rtol=1e-6, atol=1e-9)):
            bad += 1
dt = time.perf_counter() - t0
print("oracle, 1 thread: %d candidates in %.2f s -> %.0f candidate-iterations/s; %d of them differ from the GPU result beyond 1e-6"
      % (k, dt, k * 5 / dt, bad), flush=True)
pb.close()
