"""scratch: A/B of the K1 load path (LDG registers vs TMA bulk-copy pipeline) on one box"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gslnls_b200 import Model, Problem
torch.cuda.set_device(0)
NS = [int(float(v)) for v in os.environ.get("NS", "1e8,1.25e7").split(",")]
TUNES = os.environ.get("TUNES", ";".join([
    "tiled=0,block=256,unroll=4,minb=2",
    "tiled=2,block=288,unroll=2,minb=2,stages=4",
    "tiled=2,block=288,unroll=2,minb=2,stages=6",
    "tiled=2,block=288,unroll=4,minb=2,stages=3",
    "tiled=2,block=288,unroll=1,minb=3,stages=6",
    "tiled=2,block=160,unroll=2,minb=4,stages=4",
    "tiled=2,block=544,unroll=2,minb=1,stages=4",
    "tiled=2,block=544,unroll=2,minb=1,stages=6",
    "tiled=2,block=416,unroll=2,minb=1,stages=6",
])).split(";")
th = np.array([4.0, 1.3, 0.9])
for n in NS:
    x = torch.linspace(0, 3, n, dtype=torch.float64, device="cuda")
    y = 5 * torch.exp(-1.5 * x) + 1 + 0.25 * torch.randn(n, dtype=torch.float64, device="cuda")
    ref = None
    for rep in range(2):
        for tune in TUNES:
            os.environ["GSLNLS_TUNE"] = tune
            try:
                m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=False)
                pb = Problem(m, n, False, 0).bind_device([x.data_ptr()], y.data_ptr(), keepalive=(x, y))
                pk = pb.eval_packet(th)
                if ref is None:
                    ref = pk
                err = float(np.max(np.abs(pk - ref) / np.maximum(np.abs(ref), 1e-300)))
                pb.time_passes(th, 20)
                ms = min(pb.time_passes(th, 100) for _ in range(3))
                print("n=%d rep %d %-48s pass %.1f us  %.0f GB/s  max rel diff vs first %.1e" % (
                    n, rep, tune, ms * 1e3, 16.0 * n / ms / 1e6, err), flush=True)
                pb.close()
            except Exception as e:  # noqa: BLE001
                print("n=%d %-48s FAILED %s" % (n, tune, e), flush=True)
    del x, y
