#!/usr/bin/env python
"""bench_aux.py -- measurement lines for the two other GPU configurations of BASELINE.json (one B200):

    python bench_aux.py --workload gaussmix48   # configs[3]: sum of 16 Gaussians, p = 48, n = 1e7, dogleg;
                                                # the tiled FP64 DMMA pass kernel against the FP64 roofline
    python bench_aux.py --workload mstart8192   # configs[4]: 8192 multi-start candidates batched per kernel,
                                                # two-exponential mixture p = 4, n = 4096, 5 LM iterations each

`bench.py` stays the headline (configs[2]); these lines use the same keys.  Timing rules as there: warm-up,
CUDA events on the solver's stream, clocks sampled during the timed region, inputs larger than L2 or (for
mstart8192, whose 64 KB of data are meant to live in cache) stated otherwise.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_TFLOPS = 37.1   # measured DFMA = DMMA issue peak of a B200 (scratch/fp64_peak.cu, round 1)


def gaussmix_formula(K=16):
    terms, names = [], []
    for k in range(1, K + 1):
        terms.append("a%d * exp(-(x - m%d)^2 / s%d^2)" % (k, k, k))
        names += ["a%d" % k, "m%d" % k, "s%d" % k]
    return " + ".join(terms), names


def gaussmix_truth(K=16):
    th = []
    for k in range(1, K + 1):
        th += [5.0 + ((7 * k) % 11), 100.0 * (k - 0.5) / K, 2.5]
    return np.array(th)


def gaussmix_data(n, K=16, seed=4):
    """x on [0, 100], 16 Gaussians of width 2.5, noise sd 0.5"""
    th = gaussmix_truth(K)
    x = 100.0 * np.arange(n, dtype=np.float64) / float(n - 1)
    y = np.zeros(n)
    for k in range(K):
        y += th[3 * k] * np.exp(-((x - th[3 * k + 1]) ** 2) / th[3 * k + 2] ** 2)
    y += 0.5 * np.random.Generator(np.random.Philox(key=seed)).standard_normal(n)
    start = th * (1.0 + 0.02 * (-1.0) ** np.arange(3 * K))
    return x, y, start


def run_gaussmix48(args):
    import torch

    from bench import ClockSampler
    from gslnls_b200 import Model, Problem, _lib, gsl_nls_control, pack_control
    from oracle import oracle as O
    K, p, n = 16, 48, args.n or 10_000_000
    torch.cuda.set_device(0)
    x, y, start = gaussmix_data(n, K)
    x_h = torch.from_numpy(x).pin_memory()
    y_h = torch.from_numpy(y).pin_memory()
    rhs, names = gaussmix_formula(K)
    m = Model(rhs, names, ["x"], jac=True)
    pb = Problem(m, n, False, 0).upload([x_h.numpy()], y_h.numpy())
    ctrl = gsl_nls_control()

    def run_steps(k):
        left, iters, fits, last, complete = k, 0, 0, None, None
        while left > 0:
            pb.fit_begin(start, algorithm=args.algorithm, control=ctrl)
            done, run = False, 0
            while not done and run < left:
                done, r, _ = pb.fit_run(left - run)
                run += r
            last = pb.fit_end()
            left -= max(min(run, int(last["npass"])), 1)
            iters += last["niter"]
            fits += 1
            if done:
                complete = last
        return iters, fits, (complete or last)

    run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = pb.launch_count
    pb.set_profile(args.steps + 64)
    pb.timer_start()
    iters, fits, last = run_steps(args.steps)
    ms = pb.timer_stop()
    pass_ms, pass_cnt = pb.profile()
    pb.set_profile(0)
    launches = pb.launch_count - launches0
    clocks = sampler.stop()

    # e2e: the one-shot C call from pinned host buffers (H2D of x, y inside the timed region)
    import ctypes as C
    ci, cd = pack_control(ctrl, args.algorithm, False)
    L = _lib.lib()
    arr = (_lib.c_double_p * 1)(C.cast(x_h.data_ptr(), _lib.c_double_p))
    e_it, e_t = 0, 0.0
    for rep in range(3):
        res = _lib.Result()
        t0 = time.perf_counter()
        rc = L.gslnls_fit_large(m.handle, arr, C.cast(y_h.data_ptr(), _lib.c_double_p), None, n,
                                start.ctypes.data_as(_lib.c_double_p), ci.ctypes.data_as(_lib.c_int_p),
                                cd.ctypes.data_as(_lib.c_double_p), 0, 0, C.byref(res))
        dt = time.perf_counter() - t0
        _lib.check(rc)
        if rep > 0:
            e_it += res.niter
            e_t += dt
        L.gslnls_result_free(C.byref(res))

    flops = n * (p * (p + 1) + 2.0 * p + 2.0)   # J^T J (SYRK) + J^T f + f^T f per pass; the model itself not counted
    achieved = flops / (pass_ms * 1e-3) / 1e12
    value = iters / (ms * 1e-3)
    # CPU: the oracle's data flow on a row sample, one thread
    n_s = 100_000
    xs, ys, _ = gaussmix_data(n_s, K)
    t0 = time.perf_counter()
    r = O.nls_large("gaussmix", ys, start, x=xs, algorithm=args.algorithm, maxiter=10)
    cdt = time.perf_counter() - t0
    line = {
        "metric": "%s iterations/sec, gsl_nls_large, sum of 16 Gaussians n=1e7 p=48" % args.algorithm,
        "value": value, "unit": "iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "y ~ sum_k a_k exp(-(x - m_k)^2 / s_k^2), K=16, n=%d, p=48, %s "
                               "(BASELINE.json configs[3])" % (n, args.algorithm),
                   "fits": fits, "outer_iterations": iters, "l2": "inputs 160 MB exceed the 126 MB L2",
                   "final": {"ssr": float(last["ssr"]), "niter": int(last["niter"]), "status": last["status"],
                             "max_rel_err_vs_truth": float(np.max(np.abs(last["par"] / gaussmix_truth(K) - 1)))}},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": achieved / FP64_PEAK_TFLOPS, "traffic": None,
                     "kernel": "nls_pass (K1b, tiled FP64 DMMA SYRK)", "algorithmic_flops_per_launch": flops,
                     "avg_launch_ms": pass_ms, "launches_timed": int(pass_cnt),
                     "peak_source": "FP64 pipe peak measured with scratch/fp64_peak.cu (DFMA and DMMA issue to the "
                                    "same units on B200); MEASURED_PEAKS.json has no FP64 entry",
                     "note": "flops count J^T J, J^T f, f^T f only; with the 8x8 blocking and the model's 16 exp "
                             "per row the FP64 pipe needs ~0.95 ms per pass at 100 % utilisation (DESIGN.md)"},
        "e2e": {"value": e_it / e_t, "unit": "iterations/s", "h2d_bytes_per_step": int(16 * n * 2 / max(e_it, 1)),
                "d2h_bytes_per_step": 8 * (24 + 6 * p + 2 * p * p), "ms_per_fit": 1e3 * e_t / 2},
        "cpu_baseline": {"value": r["niter"] / cdt * n_s / n, "unit": "iterations/s", "cores": 1, "kind": "port",
                         "sample": "%s, first %d iterations at n=%d, 1 thread, scaled x%g"
                                   % (args.algorithm, r["niter"], n_s, n_s / n)},
    }
    print(json.dumps(line))


def run_mstart8192(args):
    import torch
    from scipy.stats import qmc

    from bench import ClockSampler
    from gslnls_b200 import Model, Problem
    from oracle import oracle as O
    n, S, iters = 4096, args.n or 8192, 5
    torch.cuda.set_device(0)
    rng = np.random.Generator(np.random.Philox(key=5))
    x = np.linspace(0, 10, n)
    y = 3 * np.exp(-0.5 * x) + 2 * np.exp(-3 * x) + 0.05 * rng.standard_normal(n)
    starts = np.ascontiguousarray(qmc.Sobol(4, scramble=False).random(S) * 10.0)
    m = Model("A1*exp(-l1*x)+A2*exp(-l2*x)", ["A1", "l1", "A2", "l2"], ["x"], jac=True)
    pb = Problem(m, n).upload([x], y)
    for _ in range(max(args.warmup, 3)):
        out = pb.fit_batch(starts, iters=iters)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = pb.launch_count
    pb.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = pb.fit_batch(starts, iters=iters)
    wall = time.perf_counter() - t0
    ms = pb.timer_stop()
    launches = pb.launch_count - launches0
    clocks = sampler.stop()
    ok = np.isfinite(out["ssr"])
    best = int(np.argmin(np.where(ok, out["ssr"], np.inf)))
    k = 64
    t0 = time.perf_counter()
    agree = 0
    for c in range(k):
        ref = O.nls_large("expmix2", y, starts[c], x=x, algorithm="lm", maxiter=iters)
        agree += int(ref["conv"] not in (0, 11) or np.allclose(out["par"][c], ref["par"], rtol=1e-6, atol=1e-9))
    cdt = time.perf_counter() - t0
    value = S * iters * args.steps / (ms * 1e-3)
    line = {
        "metric": "multi-start candidate-iterations/sec, 8192 starts x 5 LM iterations, exp mixture n=4096 p=4",
        "value": value, "unit": "candidate-iterations/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "y ~ A1*exp(-l1*x)+A2*exp(-l2*x), n=4096, p=4, S=%d Sobol starts in [0,10]^4, "
                               "mstart_p=5 LM iterations each + log det(J^T J) screen (BASELINE.json configs[4])" % S,
                   "l2": "the 64 KB of data are read by every candidate and are meant to stay in cache; no flush",
                   "step": "one batch: start upload, all passes and batched trust-region steps, result download",
                   "best": {"candidate": best, "par": [float(v) for v in out["par"][best]],
                            "ssr": float(out["ssr"][best])},
                   "oracle_agreement": "%d of %d sampled candidates within 1e-6" % (agree, k)},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": None,
        "e2e": {"value": S * iters * args.steps / wall, "unit": "candidate-iterations/s",
                "h2d_bytes_per_step": int(starts.nbytes), "d2h_bytes_per_step": int(S * 8 * 40),
                "note": "host wall clock around Problem.fit_batch() with host start / result arrays"},
        "cpu_baseline": {"value": k * iters / cdt, "unit": "candidate-iterations/s", "cores": 1, "kind": "port",
                         "sample": "%d candidates through the oracle, 1 thread" % k},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["gaussmix48", "mstart8192"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--algorithm", default="dogleg")
    ap.add_argument("--n", type=int, default=0, help="rows (gaussmix48) or candidates (mstart8192)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 40 if args.workload == "gaussmix48" else 10
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_aux.py needs a CUDA device (the product path has no CPU fallback)")
    return run_gaussmix48(args) if args.workload == "gaussmix48" else run_mstart8192(args)


if __name__ == "__main__":
    sys.exit(main())
