"""GPU parity tests of the sparse-row path (src/nls_large.c:528-648): gslnls_sparse_* through the
C ABI against the CPU oracle, which runs the same trust-region / Steihaug-Toint algorithm on the DENSE Jacobian
of the same model (what the reference does after densifying, src/nls_large.c:641-648).  Fixtures are the
reference's own: the Penalty function I of inst/unit_tests/unit_tests_gslnls.R:316-346 (p = 10) and of README
Example 4 (p = 500, SSR 0.004778845), Misra1a with a sparse Jacobian (:302-314)."""
import math

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gslnls_b200
    from gslnls_b200 import _lib
    assert _lib.lib().gslnls_device_count() > 0, "no CUDA device: the product path has no fallback"
    return gslnls_b200


def penalty_problem(G, p, alpha=1e-5, weights=None):
    """rows 0..p-1: sqrt(alpha) (theta_k - 1); row p: sum(theta^2) - 0.25 (p one-parameter terms sharing a row)"""
    sa = math.sqrt(alpha)
    idx = np.arange(p, dtype=np.int32)
    sp = G.SparseProblem(p=p, nrows=p + 1)
    sp.add_block("%.17g * (th - 1)" % sa, {"th": (0, idx)}, nterms=p)
    sp.add_block("th^2", {"th": (0, idx)}, rows=np.full(p, p, dtype=np.int32))
    y = np.zeros(p + 1)
    y[p] = 0.25
    sp.set_response(y, weights)
    return sp


def penalty_rows(p, alpha=1e-5):
    sa = math.sqrt(alpha)
    eye = np.eye(p) * sa

    def rows(th, v, wf, wJ, wh):
        f = np.concatenate([sa * (th - 1), [np.sum(th ** 2)]])
        J = np.vstack([eye, 2 * th[None, :]]) if wJ else None
        return f, J, None
    y = np.zeros(p + 1)
    y[p] = 0.25
    return rows, y


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("p", [1, 10, 500, 5000, 200_000])  # 200000: the dense row is a "long" segment (98 items)
def test_sparse_operators_penalty(G, p):
    """residual rows, J^T f and diag(J^T J) from the stored nonzeros vs the dense formulas"""
    sp = penalty_problem(G, p)
    assert sp.finalize().nnz == 2 * p
    rng = np.random.default_rng(p)
    th = rng.uniform(-1.0, 2.0, p)
    e = sp.eval(th)
    sa = math.sqrt(1e-5)
    f = np.concatenate([sa * (th - 1), [np.sum(th ** 2) - 0.25]])
    g = sa * f[:p] + 2 * th * f[p]
    d = sa * sa + 4 * th ** 2
    assert np.max(np.abs(e["resid"] - f)) <= 1e-13 * np.max(np.abs(f))
    assert np.max(np.abs(e["grad_vec"] - g)) <= 1e-12 * np.max(np.abs(g))
    assert rel(e["jtj_diag"], d) < 1e-13
    assert abs(e["ssr"] - f @ f) <= 1e-13 * (f @ f)
    # run-to-run bitwise reproducible (no atomics on data)
    e2 = sp.eval(th)
    assert np.array_equal(e["grad_vec"], e2["grad_vec"]) and e["ssr"] == e2["ssr"]
    sp.close()


@pytest.mark.parametrize("scale", ["more", "levenberg", "marquardt"])
def test_penalty_p10_unit_test_fixture(G, scale):
    """unit_tests_gslnls.R:316-346 (penalty_fit_dgC/dgR/dgT): p = 10, start 0.15, here with cgst vs the oracle"""
    p = 10
    sp = penalty_problem(G, p)
    start = np.full(p, 0.15)
    got = sp.fit(start, control={"scale": scale}, trace=True, want_jtj=True, want_resid=True)
    rows, y = penalty_rows(p)
    ref = O.nls_large(rows, y, start, algorithm="cgst", trace=True, scale=scale, want_resid_grad=True)
    assert got["conv"] == ref["conv"] == 0
    assert got["niter"] == ref["niter"] and got["info"] == ref["info"]
    assert rel(got["par"], ref["par"]) < 1e-8
    assert abs(got["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    assert rel(got["ssrtrace"], ref["ssrtrace"]) < 1e-8
    assert got["neval"]["f"] == ref["neval"]["f"] and got["neval"]["df2"] == ref["neval"]["df2"]
    assert got["neval"]["dfu"] == ref["neval"]["dfu"]
    jtj = np.tril(got["jtj"])
    assert np.max(np.abs(jtj - ref["jtj"])) <= 1e-12 * np.max(np.abs(ref["jtj"]))
    assert np.max(np.abs(got["resid"] - ref["resid"])) <= 1e-10 * np.max(np.abs(ref["resid"]))
    sp.close()


@pytest.mark.parametrize("alg", ["lm", "dogleg", "ddogleg", "subspace2D"])
def test_reference_sparse_unit_tests_with_dense_methods(G, alg, nist_problems):
    """inst/unit_tests/unit_tests_gslnls.R:302-346 run their sparse-Jacobian fits with the DEFAULT algorithm (lm):
    penalty_fit_dgC / dgR / dgT (p = 10, start 0.15) and Misra1a with a sparse Jacobian (6.2.1).  For p <= 100 the
    sparse path assembles the dense packet from its nonzeros and runs the dense trust-region kernel; vs the oracle"""
    p = 10
    sp = penalty_problem(G, p)
    start = np.full(p, 0.15)
    got = sp.fit(start, algorithm=alg, trace=True, want_jtj=True, want_resid=True)
    rows, y = penalty_rows(p)
    ref = O.nls_large(rows, y, start, algorithm=alg, trace=True, want_resid_grad=True)
    assert got["conv"] == ref["conv"] == 0 and got["niter"] == ref["niter"] and got["info"] == ref["info"]
    assert rel(got["par"], ref["par"]) < 1e-8 and abs(got["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    assert rel(got["ssrtrace"], ref["ssrtrace"]) < 1e-8
    assert got["neval"]["f"] == ref["neval"]["f"] and got["neval"]["df2"] == ref["neval"]["df2"]
    assert np.max(np.abs(np.tril(got["jtj"]) - ref["jtj"])) <= 1e-10 * np.max(np.abs(ref["jtj"]))
    assert np.max(np.abs(got["resid"] - ref["resid"])) <= 1e-10 * np.max(np.abs(ref["resid"]))
    sp.close()
    pr = nist_problems["Misra1a"]
    x, y = np.array(pr["data"]["x"]), np.array(pr["data"]["y"])
    sp = G.SparseProblem(p=2, nrows=y.size)
    sp.add_block("b1 * (1 - exp(-b2 * x))", {"b1": 0, "b2": 1}, {"x": x})
    sp.set_response(y)
    got = sp.fit(np.array(pr["start"], dtype=float), algorithm=alg)
    ref = O.nls_large(O.sympy_rows("b1 * (1 - exp(-b2 * x))", ["b1", "b2"], {"x": x}), y, pr["start"], algorithm=alg)
    assert got["conv"] == 0 and got["niter"] == ref["niter"] and rel(got["par"], ref["par"]) < 1e-8
    assert np.max(np.abs(got["par"] - np.array(pr["target"]))) < 1.22e-4   # dotest_tol("6.2.1", ...)
    sp.close()


def test_readme_example4_penalty_p500(G, readme_examples):
    """README.md:1088-1146: p = 500, start 1:p, cgst -> SSR 0.004778845 as printed; coefficients vs the oracle"""
    e = readme_examples["example4_penalty"]
    p = e["p"]
    sp = penalty_problem(G, p, alpha=e["alpha"])
    start = np.arange(1, p + 1, dtype=float)
    got = sp.fit(start, control={"maxiter": 500}, trace=True)
    assert got["conv"] == 0
    assert float("%.7g" % got["ssr"]) == e["ssr_print"]
    rows, y = penalty_rows(p, e["alpha"])
    ref = O.nls_large(rows, y, start, algorithm="cgst", maxiter=500, trace=True)
    assert got["niter"] == ref["niter"]
    assert abs(got["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    assert rel(got["ssrtrace"], ref["ssrtrace"]) < 1e-8
    # The minimum is flat (alpha = 1e-5) and the CG tolerance is 1e-6: the algorithm does not pin individual
    # coefficients to 1e-8.  Measured on the oracle itself: the same fit with the parameters listed in another
    # order -- nothing but the summation order of the length-p dot products changes -- moves them by ~7e-7
    # relative while SSR agrees to 1e-13.  The coefficient gate is that sensitivity (x 10), not 1e-8.
    perm = np.random.default_rng(0).permutation(p)
    ref2 = O.nls_large(rows, y, start[perm], algorithm="cgst", maxiter=500)
    back = np.empty(p)
    back[perm] = ref2["par"]
    order_sensitivity = rel(back, ref["par"])
    assert abs(ref2["ssr"] - ref["ssr"]) <= 1e-10 * ref["ssr"]
    assert rel(got["par"], ref["par"]) < max(1e-8, 10 * order_sensitivity)
    assert rel(got["par"], ref["par"]) < 1e-4
    assert got["launches"] == got["neval"]["f"]  # one solver launch per trial point, nothing else
    sp.close()


def test_penalty_weighted(G):
    p = 40
    w = 0.5 + (np.arange(p + 1) % 5) / 2.0
    sp = penalty_problem(G, p, weights=w)
    start = np.full(p, 0.3)
    got = sp.fit(start, trace=True)
    rows, y = penalty_rows(p)
    ref = O.nls_large(rows, y, start, weights=w, algorithm="cgst", trace=True)
    assert got["conv"] == ref["conv"] == 0 and got["niter"] == ref["niter"]
    assert rel(got["par"], ref["par"]) < 1e-8 and abs(got["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    sp.close()


def test_misra1a_sparse_jacobian(G, nist_problems):
    """unit_tests_gslnls.R:302-314 (6.2.1): Misra1a with its Jacobian returned as a sparse matrix -> certified
    values within the reference's own 1.22e-4.  Here: one block, every row one term, both parameters scalar."""
    pr = nist_problems["Misra1a"]
    x, y = np.array(pr["data"]["x"]), np.array(pr["data"]["y"])
    sp = G.SparseProblem(p=2, nrows=y.size)
    sp.add_block("b1 * (1 - exp(-b2 * x))", {"b1": 0, "b2": 1}, {"x": x})
    sp.set_response(y)
    got = sp.fit(np.array(pr["start"], dtype=float))
    assert got["conv"] == 0
    assert np.max(np.abs(got["par"] - np.array(pr["target"]))) < 1.22e-4
    rows = O.sympy_rows("b1 * (1 - exp(-b2 * x))", ["b1", "b2"], {"x": x})
    ref = O.nls_large(rows, y, pr["start"], algorithm="cgst")
    assert got["niter"] == ref["niter"] and rel(got["par"], ref["par"]) < 1e-8
    sp.close()


def grouped_problem(G, n, ngroups, seed=3):
    """y = A[g] * exp(-lam * x) + b[g]: 2 * ngroups + 1 parameters, three per row"""
    rng = np.random.Generator(np.random.Philox(key=seed))
    g = (np.arange(n) % ngroups).astype(np.int32)
    x = 3.0 * rng.random(n)
    A = 2.0 + 3.0 * rng.random(ngroups)
    b = rng.random(ngroups)
    y = A[g] * np.exp(-1.5 * x) + b[g] + 0.05 * rng.standard_normal(n)
    sp = G.SparseProblem(p=2 * ngroups + 1, nrows=n)
    sp.add_block("A * exp(-lam * x) + b", {"A": (0, g), "lam": 2 * ngroups, "b": (ngroups, g)}, {"x": x})
    sp.set_response(y)
    truth = np.concatenate([A, b, [1.5]])
    return sp, g, x, y, truth


def test_grouped_exponential_vs_oracle(G):
    """a model whose rows touch 3 of 41 parameters: fit vs the oracle on the dense n x 41 Jacobian"""
    n, ng = 6000, 20
    sp, g, x, y, truth = grouped_problem(G, n, ng)
    P = 2 * ng + 1
    start = np.concatenate([np.full(ng, 3.0), np.full(ng, 0.3), [1.0]])
    got = sp.fit(start, trace=True, want_jtj=True)

    def rows(th, v, wf, wJ, wh):
        e = np.exp(-th[2 * ng] * x)
        f = th[g] * e + th[ng + g]
        J = None
        if wJ:
            J = np.zeros((n, P))
            J[np.arange(n), g] = e
            J[np.arange(n), ng + g] = 1.0
            J[:, 2 * ng] = -th[g] * x * e
        return f, J, None
    ref = O.nls_large(rows, y, start, algorithm="cgst", trace=True)
    assert got["conv"] == ref["conv"] == 0
    assert got["niter"] == ref["niter"]
    assert rel(got["par"], ref["par"]) < 1e-8
    assert abs(got["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    assert np.max(np.abs(np.tril(got["jtj"]) - ref["jtj"])) <= 1e-12 * np.max(np.abs(ref["jtj"]))
    sp.close()


def test_grouped_exponential_vs_sparse_oracle_midsize(G):
    """n = 300000 rows, 601 parameters: the shared decay rate's column is a long segment (147 items, added by a
    warp), the group columns are consecutive runs (streamed without index lists); vs oracle/sparse.py"""
    from oracle import sparse as OS
    n, ng = 300_000, 300
    rng = np.random.Generator(np.random.Philox(key=11))
    g = (np.arange(n) // (n // ng)).astype(np.int32)
    x = 3.0 * rng.random(n)
    A = 2.0 + 3.0 * rng.random(ng)
    b = rng.random(ng)
    y = A[g] * np.exp(-1.5 * x) + b[g] + 0.05 * rng.standard_normal(n)
    w = 0.5 + rng.random(n)
    start = np.concatenate([np.full(ng, 3.0), np.full(ng, 0.3), [1.0]])
    model = OS.grouped_exp_model(g.astype(np.int64), x, ng)
    for weights in (None, w):
        sp = G.SparseProblem(p=2 * ng + 1, nrows=n)
        sp.add_block("A * exp(-lam * x) + b", {"A": (0, g), "lam": 2 * ng, "b": (ng, g)}, {"x": x})
        sp.set_response(y, weights)
        # (a) xtol = 1e-4: the step test fires while the iteration still makes progress -- everything is compared
        got = sp.fit(start, control={"xtol": 1e-4}, trace=True)
        ref = OS.nls_large_sparse(model, y, start, weights=weights, xtol=1e-4, trace=True)
        assert got["conv"] == ref["conv"] == 0 and got["niter"] == ref["niter"] and got["info"] == ref["info"]
        assert got["neval"]["f"] == ref["neval"]["f"] and got["cg_iters"] == ref["cg_iters"]
        assert got["neval"]["dfu"] == ref["neval"]["dfu"] and got["neval"]["df2"] == ref["neval"]["df2"]
        assert rel(got["par"], ref["par"]) < 1e-8
        assert abs(got["ssr"] - ref["ssr"]) <= 1e-10 * ref["ssr"]
        assert rel(got["ssrtrace"], ref["ssrtrace"]) < 1e-8
        assert np.max(np.abs(got["grad_vec"] - ref["grad_vec"])) <= 1e-8 * max(1.0, np.max(np.abs(ref["grad_vec"])))
        # (b) default xtol = 1.5e-8: the last iterations make no progress in double precision (the oracle's trace:
        # SSR constant to the last bit over its iterations 5..7, every trial step accepted or rejected on two norms
        # that agree to 1e-16, i.e. on rounding), so the iteration at which the step test fires -- and with it the
        # last ~10 xtol of the coefficients -- is not defined by the algorithm.  Compared: the common part of the
        # trace, SSR, coefficients to 1e-6.
        got = sp.fit(start, trace=True)
        ref = OS.nls_large_sparse(model, y, start, weights=weights, trace=True)
        assert got["conv"] == ref["conv"] == 0 and abs(got["niter"] - ref["niter"]) <= 2
        m = min(got["niter"], ref["niter"]) + 1
        assert rel(got["ssrtrace"][:m], ref["ssrtrace"][:m]) < 1e-8
        assert abs(got["ssr"] - ref["ssr"]) <= 1e-10 * ref["ssr"]
        assert rel(got["par"], ref["par"]) < 1e-6
        sp.close()


def test_grouped_exponential_large_properties(G):
    """n = 4e6 rows, 20001 parameters (a dense J would be 640 GB): the fit converges to the generating
    parameters, the gradient at the solution vanishes, and a second run is bitwise identical"""
    n, ng = 4_000_000, 10_000
    sp, g, x, y, truth = grouped_problem(G, n, ng)
    start = np.concatenate([np.full(ng, 3.0), np.full(ng, 0.3), [1.0]])
    a = sp.fit(start)
    assert a["conv"] == 0 and a["nnz"] == 3 * n
    assert abs(a["par"][-1] - 1.5) < 5e-3
    assert np.max(np.abs(a["par"][:ng] - truth[:ng])) < 0.2
    assert np.max(np.abs(a["grad_vec"])) < 1e-5 * a["ssr"]
    e = sp.eval(a["par"])
    assert abs(e["ssr"] - a["ssr"]) <= 1e-12 * a["ssr"]
    b = sp.fit(start)
    assert np.array_equal(a["par"], b["par"]) and a["ssr"] == b["ssr"] and a["niter"] == b["niter"]
    sp.close()


def test_sparse_error_paths(G):
    sp = penalty_problem(G, 101)                     # the dense trust-region kernel stops at p = 100
    with pytest.raises(Exception, match="cgst"):
        sp.fit(np.full(101, 0.15), algorithm="lm")
    sp.close()
    sp = penalty_problem(G, 4)
    with pytest.raises(Exception, match="fvv"):
        sp.fit(np.full(4, 0.15), algorithm="lmaccel")
    sp.close()
    # fewer rows than parameters (R/nls_large.R: negative residual degrees of freedom)
    sp = G.SparseProblem(p=3, nrows=2)
    sp.add_block("a * x", {"a": (0, np.array([0, 1], dtype=np.int32))}, {"x": np.array([1.0, 2.0])})
    with pytest.raises(Exception, match="degrees of freedom"):
        sp.fit(np.zeros(3))
    sp.close()
    # parameter index out of range is refused at finalize
    sp = G.SparseProblem(p=2, nrows=2)
    sp.add_block("a * x", {"a": (0, np.array([0, 2], dtype=np.int32))}, {"x": np.array([1.0, 2.0])})
    with pytest.raises(Exception, match="out of range"):
        sp.finalize()
    sp.close()
    # a non-finite Jacobian entry -> GSL_EBADFUNC, start values returned (src/nls_large.c:560-566, :293-302)
    sp = G.SparseProblem(p=2, nrows=3)
    sp.add_block("a * log(b * x)", {"a": 0, "b": 1}, {"x": np.array([1.0, 2.0, 0.0])})
    sp.set_response(np.ones(3))
    from gslnls_b200 import _lib
    res = None
    try:
        res = sp.fit(np.array([1.0, 1.0]))
    except _lib.GslnlsError as err:
        assert "function" in str(err).lower() or "bad" in str(err).lower()
    if res is not None:
        assert res["conv"] != 0 and np.array_equal(res["par"], [1.0, 1.0])
    sp.close()
