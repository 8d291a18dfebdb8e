"""CPU-side checks of the sparse-row path: the term-evaluation kernel is part of every model
with a symbolic Jacobian and at most 16 parameters and compiles for sm_100a; the entry points refuse to run
without a device (no CPU path); the oracle side of the GPU parity tests (dense Jacobian + cgst) reproduces the
reference's fixtures."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from gslnls_b200 import _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("rhs,names,nvar,expect", [
    ("th^2", ["th"], 0, True),                                   # Penalty I, the dense row as p terms
    ("A * exp(-lam * x) + b", ["A", "lam", "b"], 1, True),        # grouped exponential
    ("+".join("a%d*x^%d" % (i, i) for i in range(17)), ["a%d" % i for i in range(17)], 1, False),  # > 16 slots
])
def test_sparse_eval_kernel_compiles_for_sm100a(tmp_path, rhs, names, nvar, expect):
    cub = str(tmp_path / "m.cubin")
    code = ("import sys; sys.path.insert(0, %r)\nfrom gslnls_b200 import Model\n"
            "Model(%r, %r, %r, jac='symbolic')\n" % (ROOT, rhs, names, ["x"][:nvar]))
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GSLNLS_DUMP_CUBIN=cub), check=True)
    out = subprocess.run(["cuobjdump", "-elf", cub], capture_output=True, text=True, check=True).stdout
    assert ("nls_sparse_eval" in out) == expect


@pytest.mark.skipif(_lib.lib().gslnls_device_count() > 0, reason="a GPU is present")
def test_sparse_entry_points_refuse_without_a_device():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.gslnls_sparse_create(0, 10, 11, C.byref(h)) == 1004 and not h.value
    assert b"no CPU path" in L.gslnls_last_error()
    # argument validation comes first
    assert L.gslnls_sparse_create(0, 0, 11, C.byref(h)) == 4
    assert L.gslnls_sparse_fit(None, None, None, None, 0, 0, None) == 4
    assert L.gslnls_sparse_finalize(None) == 4
    assert L.gslnls_sparse_nnz(None) == 0
    L.gslnls_sparse_free(None)
    from gslnls_b200 import SparseProblem
    with pytest.raises(_lib.GslnlsError):
        SparseProblem(p=10, nrows=11)


def test_oracle_side_of_the_sparse_parity_tests():
    """unit_tests_gslnls.R:316-346: Penalty I, p = 10, start 0.15 -- all four Jacobian classes give one fit; the
    oracle's cgst on the dense Jacobian converges to the same minimum as its lm (the reference's default)"""
    p, sa = 10, math.sqrt(1e-5)
    eye = np.eye(p) * sa

    def rows(th, v, wf, wJ, wh):
        return (np.concatenate([sa * (th - 1), [np.sum(th ** 2) - 0.25]]),
                np.vstack([eye, 2 * th[None, :]]) if wJ else None, None)
    a = O.nls_large(rows, np.zeros(p + 1), np.full(p, 0.15), algorithm="cgst")
    b = O.nls_large(rows, np.zeros(p + 1), np.full(p, 0.15), algorithm="lm")
    assert a["conv"] == b["conv"] == 0
    assert np.allclose(a["par"], b["par"], rtol=1e-5) and abs(a["ssr"] - b["ssr"]) < 1e-9
    # More', Garbow, Hillstrom (1981) problem 23, n = 10: f* = 7.08765e-5
    assert abs(a["ssr"] - 7.08765e-5) < 1e-9


# ------------------------------------------------------------------- oracle/sparse.py pinned on the dense oracle
def _dense_rows_from(model, n, p):
    def rows(th, v, wf, wJ, wh):
        f, trip = model(th, wJ)
        J = None
        if wJ:
            J = np.zeros((n, p))
            np.add.at(J, (trip[0], trip[1]), trip[2])
        return f, J, None
    return rows


@pytest.mark.parametrize("p,scale,weighted", [(10, "more", False), (10, "levenberg", False), (10, "marquardt", False),
                                               (40, "more", True), (500, "more", False)])
def test_sparse_oracle_matches_dense_oracle_penalty(p, scale, weighted):
    from oracle import sparse as OS
    model, y = OS.penalty_model(p)
    start = np.full(p, 0.15) if p < 500 else np.arange(1, p + 1, dtype=float)
    w = (0.5 + (np.arange(p + 1) % 5) / 2.0) if weighted else None
    a = OS.nls_large_sparse(model, y, start, weights=w, scale=scale, maxiter=500, trace=True)
    b = O.nls_large(_dense_rows_from(model, p + 1, p), y, start, weights=w, algorithm="cgst", scale=scale,
                    maxiter=500, trace=True)
    assert a["conv"] == b["conv"] == 0 and a["niter"] == b["niter"] and a["info"] == b["info"]
    assert a["neval"] == b["neval"]
    assert abs(a["ssr"] - b["ssr"]) <= 1e-10 * b["ssr"]
    assert np.max(np.abs(a["ssrtrace"] - b["ssrtrace"]) / b["ssrtrace"]) < 1e-9
    # p = 500: flat minimum, coefficients defined to ~1e-6 only (see tests/test_gpu_sparse.py)
    assert np.max(np.abs(a["par"] - b["par"]) / np.abs(b["par"])) < (1e-8 if p < 500 else 1e-5)
    if p == 500:
        assert float("%.7g" % a["ssr"]) == 0.004778845  # README.md:1100


def test_sparse_oracle_matches_dense_oracle_grouped():
    from oracle import sparse as OS
    n, ng = 3000, 12
    rng = np.random.default_rng(5)
    g = (np.arange(n) * ng // n).astype(np.int64)
    x = 3.0 * rng.random(n)
    y = (2.0 + g % 3)[...] * np.exp(-1.5 * x) + 0.1 * g + 0.05 * rng.standard_normal(n)
    model = OS.grouped_exp_model(g, x, ng)
    start = np.concatenate([np.full(ng, 3.0), np.full(ng, 0.3), [1.0]])
    a = OS.nls_large_sparse(model, y, start)
    b = O.nls_large(_dense_rows_from(model, n, 2 * ng + 1), y, start, algorithm="cgst")
    assert a["conv"] == b["conv"] == 0 and a["niter"] == b["niter"] and a["neval"] == b["neval"]
    assert np.max(np.abs(a["par"] - b["par"]) / np.abs(b["par"])) < 1e-9
    assert abs(a["ssr"] - b["ssr"]) <= 1e-11 * b["ssr"]


def test_threaded_gather_list_builder_matches_the_serial_one():
    """gslnls_b200/csrc/seg_build.hpp (host side of sp_finalize): the threaded stable grouping reproduces the serial
    counting sort entry for entry on random / sorted / interleaved keys, items and their classes follow their
    definitions -- compiled and run on the host by tests/host_harness/seg_main.cpp"""
    d = os.path.join(ROOT, "tests", "host_harness")
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", d, "-s", "_build/seg_main"], env=env)
    r = subprocess.run([os.path.join(d, "_build", "seg_main")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "seg_build: ok" in r.stdout, r.stdout + r.stderr
