"""scratch: does asking the head of the shard to stay in L2 (evict_last) pay from pass to pass?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gslnls_b200 import Model, Problem
torch.cuda.set_device(0)

def ev_time(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

# what this box's HBM does: device-to-device copy (read + write) and a pure read
n0 = 100_000_000
src = torch.empty(2 * n0, dtype=torch.float64, device="cuda").normal_()
dst = torch.empty_like(src)
ms = min(ev_time(lambda: dst.copy_(src), 10) for _ in range(3))
print("box: copy 1.6 GB -> 1.6 GB: %.1f us, %.0f GB/s (read+write)" % (ms * 1e3, 2 * src.numel() * 8 / ms / 1e6), flush=True)
ms = min(ev_time(lambda: torch.sum(src), 10) for _ in range(3))
print("box: torch.sum over 1.6 GB: %.1f us, %.0f GB/s (read)" % (ms * 1e3, src.numel() * 8 / ms / 1e6), flush=True)
del src, dst

NS = [int(float(v)) for v in os.environ.get("NS", "1e8,5e7,2.5e7,1.25e7").split(",")]
KEEPS = [float(v) for v in os.environ.get("KEEPS", "0,0.001,32,64,80,96,112").split(",")]
TUNES = os.environ.get("TUNES", "tiled=0,block=256,unroll=4,minb=2;tiled=2,block=416,unroll=2,minb=1,stages=6").split(";")
th = np.array([4.0, 1.3, 0.9])
for n in NS:
    x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
    y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
    ref = None
    for tune in TUNES:
        for keep in KEEPS:
            os.environ["GSLNLS_TUNE"] = tune
            os.environ["GSLNLS_L2_KEEP_MB"] = str(keep)
            try:
                m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=False)
                pb = Problem(m, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
                pk = pb.eval_packet(th)
                if ref is None:
                    ref = pk
                err = float(np.max(np.abs(pk - ref) / np.maximum(np.abs(ref), 1e-300)))
                pb.time_passes(th, 20)
                ms = min(pb.time_passes(th, 100) for _ in range(3))
                print("n=%d %-46s keep %6.3f MB  pass %.1f us  %.0f GB/s algorithmic  rel diff %.1e" % (
                    n, tune, keep, ms * 1e3, 16.0 * n / ms / 1e6, err), flush=True)
                pb.close()
            except Exception as e:  # noqa: BLE001
                print("n=%d %-46s keep %s FAILED %s" % (n, tune, keep, e), flush=True)
    del x, y
