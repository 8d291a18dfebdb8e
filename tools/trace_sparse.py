#!/usr/bin/env python
"""Per-phase times of the sparse solver's CG iteration (GSLNLS_SP_TRACE stamps written by CTA 0 after each
grid.sync of the first ~60 CG iterations of every launch) on the bench workload.  Usage: tools/trace_sparse.py [n]"""
import collections
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
path = os.path.join(tempfile.gettempdir(), "sp_trace.txt")
from gslnls_b200 import SparseProblem  # noqa: E402

g, x, y, ng, start = bench.sparse_rows(n)
sp = SparseProblem(p=2 * ng + 1, nrows=n)
sp.add_block("A * exp(-lam * x) + b", {"A": (0, g), "lam": 2 * ng, "b": (ng, g)}, {"x": x})
sp.set_response(y).finalize()
sp.fit(start)
if os.path.exists(path):
    os.remove(path)
os.environ["GSLNLS_SP_TRACE"] = path
r = sp.fit(start)
del os.environ["GSLNLS_SP_TRACE"]
names = {(0, 1): "launch start -> first CG prologue (trial-point work)", (1, 2): "J d (term products)",
         (2, 3): "z update + column items (J^T u)", (3, 4): "column totals + r update", (4, 1): "d update + P-vector prologue"}
dur = collections.defaultdict(list)
prev = None
for ln in open(path):
    if ln.startswith("launch"):
        prev = None
        continue
    code, t = (int(v) for v in ln.split())
    if prev is not None:
        dur[(prev[0], code)].append((t - prev[1]) / 1e3)
    prev = (code, t)
print("n=%d  fit: %d CG iterations, solver %.1f ms" % (n, r["cg_iters"], r["solver_ms"]))
tot = 0.0
for k in [(1, 2), (2, 3), (3, 4), (4, 1), (0, 1)]:
    v = np.array(dur.get(k, [0.0]))
    print("%-52s median %8.1f us  mean %8.1f  (%d samples)" % (names[k], np.median(v), v.mean(), v.size))
    if k != (0, 1):
        tot += np.median(v)
print("CG iteration (sum of medians): %.1f us" % tot)
