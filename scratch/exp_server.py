"""scratch experiment: where does a trust-region iteration spend its time (server mode)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_rows, START, FORMULA_RHS
from gslnls_b200 import Model, Problem, gsl_nls_control

n = int(float(os.environ.get("N", "1e8")))
x, y = synth_rows(0, n, n)
model = Model(FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
pb = Problem(model, n, False, 0).upload([x], y)
ctrl = gsl_nls_control()
for alg in ("lm", "lmaccel", "dogleg"):
    for rep in range(3):
        pb.channel_stats(reset=True)
        pb.timer_start()
        t0 = time.perf_counter()
        r = pb.fit(np.array(START), algorithm=alg, control=ctrl)
        wall = time.perf_counter() - t0
        ms = pb.timer_stop()
        a, b, k = pb.channel_stats()
        print(alg, "rep", rep, "niter", r["niter"], "npass", r["npass"], "device ms %.3f wall ms %.3f" % (ms, wall * 1e3),
              "per pass %.1f us" % (1e3 * ms / max(r["npass"], 1)), "stream %.1f us step %.1f us passes %d" % (a, b, k), flush=True)
print("time_passes (launch-ordered, back to back): %.1f us" % (1e3 * pb.time_passes(np.array(START), 20)))
