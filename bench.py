#!/usr/bin/env python
"""bench.py -- LM iterations/sec of gsl_nls_large(method='lm') on the synthetic exponential model
y ~ A*exp(-lam*x)+b, n = 1e8, p = 3 (BASELINE.json configs[2]), observation-sharded over N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one trust-region trial iteration: one fused pass over all n observations (residuals,
Jacobian rows, J^T J, J^T f, f^T f) followed by the device-side trust-region step.  Whole fits are
run back to back from the README start values until exactly K steps have been executed; `value` is
the number of LM (outer) iterations those steps completed per second of device time, with the data
resident in HBM.  `e2e` is the same metric through gslnls_fit_large() with HOST buffers: pinned-host
to device copies of x and y, the fit, and the result read-back all inside the timed region.
Rank 0 prints one JSON line.
"""
import argparse
import json
import os

# NCCL writes its version / INFO lines to stdout by default; the contract is ONE JSON line there
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
# CUDA-event pairs around every 4th pass launch inside the timed region give roofline.avg_launch_ms; an event
# record between every two launches would cost each step a few microseconds of stream serialisation
os.environ.setdefault("GSLNLS_PROF_STRIDE", "4")

import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 100_000_000
CHUNK = 1 << 20
TRUTH = (5.0, 1.5, 1.0)
START = (1.0, 1.0, 0.0)   # well-conditioned start; README's (0,0,0) start is a parity case
FORMULA_RHS = "A * exp(-lam * x) + b"


def synth_rows(lo, hi, n_total, seed=1):
    """rows [lo, hi) of the synthetic data set; any shard regenerates identical doubles
    (counter-based Philox stream keyed by (seed, chunk index))"""
    x = 3.0 * np.arange(lo, hi, dtype=np.float64) / float(n_total - 1)
    z = np.empty(hi - lo)
    c0, c1 = lo // CHUNK, (hi - 1) // CHUNK
    for c in range(c0, c1 + 1):
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        g = np.random.Generator(np.random.Philox(key=[seed, c]))
        full = g.standard_normal(CHUNK)
        z[a - lo:b - lo] = full[a - c * CHUNK:b - c * CHUNK]
    y = TRUTH[0] * np.exp(-TRUTH[1] * x) + TRUTH[2] + 0.25 * z
    return x, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per K1 launch from the committed ncu --set full capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as fh:
            return json.load(fh)
    except Exception:  # noqa: BLE001
        return None


def cpu_fit(n_sample, threads, n_total):
    """the reference's CPU data flow (oracle port of src/nls_large.c + GSL multilarge) on a sample"""
    from oracle import oracle as O
    # same design (x grid on [0,3], same noise law) at reduced n, so that the trajectory matches the full fit's
    x, y = synth_rows(0, n_sample, n_sample)
    t0 = time.perf_counter()
    r = O.nls_large("exp3", y, START, x=x, algorithm="lm", threads=threads)
    dt = time.perf_counter() - t0
    return r, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_s = args.cpu_sample
    from oracle import oracle as O
    O.build()
    iters, elapsed, last = 0, 0.0, None
    for _ in range(max(args.warmup, 0) and 1):
        cpu_fit(min(n_s, 1_000_000), cores, N_FULL)
    while iters < args.steps:
        r, dt = cpu_fit(n_s, cores, N_FULL)
        iters += r["niter"]
        elapsed += dt
        last = r
    scale = n_s / float(N_FULL)
    value = iters / elapsed * scale
    line = {
        "impl": "reference", "metric": "LM iterations/sec, gsl_nls_large(lm), exp model n=1e8 p=3",
        "value": value, "unit": "iterations/s", "n_gpus": args.gpus, "steps": iters, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "y ~ A*exp(-lam*x)+b, n=1e8, p=3, lm, scale=more (BASELINE.json configs[2])",
                   "note": "reference CPU algorithm (oracle port: src/nls_large.c data flow + GSL multilarge "
                           "restatement; R and libgsl are not installable here); timed on a row sample and "
                           "scaled by n_sample/n (the path is O(n) per iteration)"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port",
                         "sample": "same design at n=%d, %d full lm fits (%d iterations), scaled x%g" % (
                             n_s, max(1, iters // max(last["niter"], 1)), iters, scale)},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL)
    ap.add_argument("--algorithm", default="lm")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000)
    ap.add_argument("--e2e-fits", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from gslnls_b200 import Model, Problem, gsl_nls_control, pack_control
    from gslnls_b200 import _lib
    from gslnls_b200.distributed import init_comm_from_torch, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    lo, hi = shard_bounds(n, rank, world)
    n_loc = hi - lo

    # ---- synthetic shard: pinned host buffers (for the e2e leg) and a resident device copy --------
    x_h = torch.empty(n_loc, dtype=torch.float64).pin_memory()
    y_h = torch.empty(n_loc, dtype=torch.float64).pin_memory()
    xs, ys = synth_rows(lo, hi, n)
    x_h.numpy()[:] = xs
    y_h.numpy()[:] = ys
    del xs, ys
    model = Model(FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    comm = init_comm_from_torch(local) if world > 1 else None
    pb = Problem(model, n_loc, False, local).upload([x_h.numpy()], y_h.numpy())
    if comm is not None:
        pb.set_comm(comm)
    ctrl = gsl_nls_control()
    start = np.array(START)

    def run_steps(k):
        """exactly k trial steps as back-to-back fits; returns (outer iterations, fits, last result)"""
        left, iters, fits, last, complete = k, 0, 0, None, None
        while left > 0:
            t0 = time.perf_counter()
            pb.fit_begin(start, algorithm=args.algorithm, control=ctrl)
            t1 = time.perf_counter()
            done, run = False, 0
            while not done and run < left:
                done, r, _ = pb.fit_run(left - run)
                run += r
            t2 = time.perf_counter()
            last = pb.fit_end()
            t3 = time.perf_counter()
            host_us[0] += (t1 - t0) * 1e6
            host_us[1] += (t2 - t1) * 1e6
            host_us[2] += (t3 - t2) * 1e6
            host_us[3] += 1
            # trailing no-op launches after convergence are not steps
            used = min(run, int(last["npass"]))
            left -= max(used, 1)
            iters += last["niter"]
            fits += 1
            if done:
                complete = last
        return iters, fits, (complete or last)

    host_us = [0.0, 0.0, 0.0, 0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pb.launch_count
    pb.set_profile(args.steps + 64)
    pb.channel_stats(reset=True)
    barrier()
    host_us[:] = [0.0, 0.0, 0.0, 0]
    pb.timer_start()
    iters, fits, last = run_steps(args.steps)
    ms = pb.timer_stop()
    barrier()
    pass_ms, pass_cnt = pb.profile()
    stream_us, step_us, stream_cnt = pb.channel_stats()
    pb.set_profile(0)
    launches = pb.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = iters / (ms * 1e-3)

    # ---- e2e: gslnls_fit_large[_sharded]() from pinned host buffers, copies inside the timed region ----
    # every rank passes its shard as HOST pointers; the call uploads it over that GPU's PCIe link, fits
    # (packets cross NVLink) and returns the result record.  Host wall clock around the call between
    # barriers (the call is synchronous and includes host work), max over ranks.
    e2e = None
    if args.e2e_fits > 0:
        import ctypes as C
        ci, cd = pack_control(ctrl, args.algorithm, False)
        L = _lib.lib()
        arr = (_lib.c_double_p * 1)(C.cast(x_h.data_ptr(), _lib.c_double_p))
        yp = C.cast(y_h.data_ptr(), _lib.c_double_p)
        e_iters, e_t, d2h, e_last = 0, 0.0, 0, None
        for rep in range(args.e2e_fits + 1):
            res = _lib.Result()
            barrier()
            t0 = time.perf_counter()
            rc = L.gslnls_fit_large_sharded(model.handle, arr, yp, None, n_loc,
                                            start.ctypes.data_as(_lib.c_double_p),
                                            ci.ctypes.data_as(_lib.c_int_p), cd.ctypes.data_as(_lib.c_double_p),
                                            local, comm.handle if comm is not None else None, 0, C.byref(res))
            dt = time.perf_counter() - t0
            _lib.check(rc)
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if rep > 0:  # the first call loads the kernels and fills the library's buffer cache
                e_iters += res.niter
                e_t += dt
                d2h += 8 * (24 + 6 * 3 + 2 * 9) * world
                e_last = [res.par[i] for i in range(3)]
            L.gslnls_result_free(C.byref(res))
        e2e = {"value": e_iters / e_t, "unit": "iterations/s",
               "h2d_bytes_per_step": int(16 * n * args.e2e_fits / max(e_iters, 1)),
               "d2h_bytes_per_step": int(d2h / max(e_iters, 1)),
               "ms_per_fit": 1e3 * e_t / args.e2e_fits, "par": e_last,
               "note": "gslnls_fit_large_sharded() per rank: pinned-host H2D of the rank's x,y shard + full lm fit "
                       "to convergence + result D2H per call; %d calls, %d iterations; wall clock between "
                       "barriers, max over ranks; bytes are totals over all ranks per outer iteration"
                       % (args.e2e_fits, e_iters)}

    # the sampler ran through both timed regions (device-resident steps and the e2e calls)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return 0

    peak, peak_src = measured_peaks()
    alg_bytes = 16.0 * n_loc  # 8 B x (1 predictor + 1 response) per observation, nothing O(n) written
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else None
    traffic = recorded_traffic()
    line = {
        "metric": "LM iterations/sec, gsl_nls_large(%s), exp model n=1e8 p=3" % args.algorithm,
        "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "y ~ A*exp(-lam*x)+b, n=%d, p=3, %s, scale=more, start=(1,1,0) "
                               "(BASELINE.json configs[2])" % (n, args.algorithm),
                   "n_per_gpu": n_loc, "fits": fits, "outer_iterations": iters,
                   "passes_per_iteration": args.steps / max(iters, 1),
                   "host_us_per_fit": {"fit_begin": host_us[0] / max(host_us[3], 1),
                                       "fit_run": host_us[1] / max(host_us[3], 1),
                                       "fit_end": host_us[2] / max(host_us[3], 1)},
                   "l2": "inputs (%.2f GB per GPU) exceed the 126 MB L2; no flush needed" % (alg_bytes / 1e9),
                   "final": {"par": [float(v) for v in last["par"]], "ssr": float(last["ssr"]),
                             "niter": int(last["niter"]), "status": last["status"]},
                   "parallelism": "observation-sharded x%d; per pass one %d-double packet per rank, %s" % (
                       world, 3 * 4 // 2 + 3 + 2,
                       "deposited by the pass kernel in every GPU's mailbox over NVLink peer memory and summed in "
                       "rank order by the resident trust-region warp (no collective call)"
                       if (comm is not None and comm.has_peer_memory) else
                       ("NCCL all-gather + rank-order sum" if world > 1 else "single GPU, resident trust-region warp"))},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None,
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "kernel": "nls_pass (K1)", "algorithmic_bytes_per_launch": alg_bytes,
                     "avg_launch_ms": pass_ms, "launches_timed": int(pass_cnt),
                     "launch_sampling": "every %s-th pass launch of the timed region" % os.environ["GSLNLS_PROF_STRIDE"],
                     "peak_source": peak_src,
                     "note": "avg_launch_ms = CUDA-event time of the nls_pass launches inside the timed region; in "
                             "resident-server mode a launch starts by waiting (in-kernel) for the trust-region warp's "
                             "request, so it spans wait + stream + grid reduction; stream_us / step_us are the device "
                             "globaltimer splits (request seen -> packet out, packet complete -> next request)",
                     "stream_us": stream_us, "step_us": step_us,
                     "frac_stream": (alg_bytes / (stream_us * 1e-6) / 1e9 / peak) if stream_us > 0 else None},
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        n_s = args.cpu_sample
        r, dt = cpu_fit(n_s, 1, N_FULL)
        scale = n_s / float(N_FULL)
        line["cpu_baseline"] = {"value": r["niter"] / dt * scale, "unit": "iterations/s", "cores": 1, "kind": "port",
                                "sample": "one full lm fit (%d iterations) on the same design at n=%d, 1 thread "
                                          "(the reference is single-threaded), scaled x%g; omits the R interpreter"
                                          % (r["niter"], n_s, scale)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
