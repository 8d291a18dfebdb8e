"""Pins of the CPU oracle (oracle/) against every number the reference records for the
trust-region algorithm: README.md traces / counts / printed answers and NIST certified values.
CPU-only; these are the checks that let the GPU parity tests trust the oracle."""
import math

import numpy as np
import pytest

from oracle import oracle as O


def _sig(v, nd=6):
    return float("%.*g" % (nd, v))


@pytest.fixture(scope="module")
def ex2(readme_examples):
    e = readme_examples["example2"]
    return e, np.array(e["x"]), np.array(e["y"])


def test_example1_lm_matches_readme(readme_examples):
    # README.md:178-195, 246-261: 9 iterations, coefficients, SSR, standard errors
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    for fd in (0, 1):
        r = O.nls_large("exp3", y, e["start"], x=x, algorithm="lm", fd_jac=fd)
        assert r["conv"] == 0 and r["niter"] == e["niter"]
        assert [round(v, 3) for v in r["par"]] == e["coef_print"]
        assert round(r["ssr"], 3) == e["ssr_print"]
        se = np.sqrt(np.diag(r["covar"]) * r["ssr"] / (len(y) - 3))
        assert [round(v, 4) for v in se] == e["stderr_print"]
        assert abs(r["ssrtol"]) < 1e-14


def test_example2_lm_trace(ex2):
    # README.md:568-605: multifit lm with forward-difference Jacobian, 26 iterations
    e, x, y = ex2
    r = O.nls_large("gauss", y, e["start"], x=x, algorithm="lm", fd_jac=1, trace=True)
    g = e["lm_fd"]
    assert r["niter"] == g["niter"] and r["conv"] == 0
    assert _sig(r["chisq_init"]) == e["ssr_init_print"]
    # 124 in the reference; the count depends on rounding-level accept/reject decisions
    assert abs(r["neval"]["f"] - g["nevalf"]) <= 1
    for t in g["trace"]:
        k = t["iter"]
        # iterations 16-19 pass through a region where 1e-8 Jacobian noise is amplified
        tol = 1e-4 if 16 <= k <= 19 else 6e-6  # 6 printed digits
        assert r["ssrtrace"][k] == pytest.approx(t["ssr"], rel=tol), k
        assert np.allclose(r["partrace"][k], t["par"], rtol=tol, atol=1e-6), k


@pytest.mark.parametrize("key,fd_fvv", [("lmaccel_fd", 1), ("lmaccel_fvv", 0)])
def test_example2_lmaccel_trace(ex2, key, fd_fvv):
    # README.md:636-659 (FD fvv) and :772-804 (analytic fvv): 12 iterations, every printed digit
    e, x, y = ex2
    r = O.nls_large("gauss", y, e["start"], x=x, algorithm="lmaccel", fd_jac=1, fd_fvv=fd_fvv, trace=True)
    g = e[key]
    assert r["niter"] == g["niter"] and r["conv"] == 0
    if "nevalf" in g:
        assert r["neval"]["f"] == g["nevalf"]  # 76 = 4 + 12*3 + 18 trials * 2
    for t in g["trace"]:
        k = t["iter"]
        assert r["ssrtrace"][k] == pytest.approx(t["ssr"], rel=6e-6), k
        assert np.allclose(r["partrace"][k], t["par"], rtol=6e-6, atol=1e-6), k
    assert [round(v, 4) for v in r["par"]] == e["coef_print"]


def _branin():
    a = [-5.1 / (4 * math.pi ** 2), 5 / math.pi, -6, 10, 1 / (8 * math.pi)]

    def rows(th, v, wf, wJ, wh):
        x1, x2 = th
        f = np.array([x2 + a[0] * x1 ** 2 + a[1] * x1 + a[2], math.sqrt(a[3] * (1 + (1 - a[4]) * math.cos(x1)))])
        J = np.array([[2 * a[0] * x1 + a[1], 1.0], [-a[3] * (1 - a[4]) * math.sin(x1) / (2 * f[1]), 0.0]])
        h = None
        if wh:
            g = a[3] * (1 + (1 - a[4]) * math.cos(x1))
            gp = -a[3] * (1 - a[4]) * math.sin(x1)
            gpp = -a[3] * (1 - a[4]) * math.cos(x1)
            h = np.array([2 * a[0] * v[0] ** 2, (gpp / (2 * math.sqrt(g)) - gp * gp / (4 * g ** 1.5)) * v[0] ** 2])
        return f, J, h
    return rows


def test_example3_branin_minima(readme_examples):
    # README.md:925-975: lm ends at (-pi, 12.275) after 20 iterations, all other methods at (pi, 2.275)
    e = readme_examples["example3_branin"]
    rows = _branin()
    r = O.nls_large(rows, np.zeros(2), e["start"], algorithm="lm", fd_jac=1)
    assert r["niter"] == e["lm_niter"]
    assert [round(v, 3) for v in r["par"]] == e["lm_coef_print"]
    assert round(r["ssr"], 4) == e["ssr_print"]
    for alg in ("lmaccel", "dogleg", "ddogleg", "subspace2D"):
        r = O.nls_large(rows, np.zeros(2), e["start"], algorithm=alg, fd_jac=1, fd_fvv=int(alg == "lmaccel"))
        assert np.allclose(r["par"], e["other_methods_min"], atol=1e-6), alg


def test_example4_penalty_cgst(readme_examples):
    # README.md:1088-1101: gsl_nls_large(cgst) on the p=500 penalty function, SSR 0.004778845
    e = readme_examples["example4_penalty"]
    p = e["p"]
    sa = math.sqrt(e["alpha"])
    eye = np.eye(p) * sa

    def rows(th, v, wf, wJ, wh):
        f = np.concatenate([sa * (th - 1), [np.sum(th ** 2) - 0.25]])
        J = np.vstack([eye, 2 * th[None, :]]) if wJ else None
        return f, J, None
    r = O.nls_large(rows, np.zeros(p + 1), np.arange(1, p + 1, dtype=float), algorithm="cgst", maxiter=500)
    assert r["conv"] == 0
    assert float("%.7g" % r["ssr"]) == e["ssr_print"]


# problems every solver is expected to solve from NIST "start 1" (the reference's own unit tests
# use Misra1a with lm/dogleg/lmaccel at 1.22e-4 absolute, inst/unit_tests/unit_tests_gslnls.R:108-115)
EASY = ["Misra1a", "Chwirut2", "Chwirut1", "Gauss1", "Gauss2", "DanWood", "Misra1b", "Kirby2", "Hahn1",
        "Gauss3", "Misra1c", "Misra1d", "Roszman1", "Thurber", "Ratkowsky2", "ENSO"]


@pytest.mark.parametrize("name", EASY)
def test_nist_certified_values(nist_problems, name):
    pr = nist_problems[name]
    lhs, rhs = O.split_formula(pr["formula"])
    data = {k: np.array(v) for k, v in pr["data"].items()}
    rows = O.sympy_rows(rhs, pr["param_names"], {k: v for k, v in data.items() if k != "y"})
    y = data["y"]
    for alg in ("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"):
        r = O.nls_large(rows, y, pr["start"], algorithm=alg)
        assert r["conv"] == 0, (name, alg)
        rel = np.max(np.abs(r["par"] - np.array(pr["target"])) / np.abs(pr["target"]))
        assert rel < 1e-6, (name, alg, rel)


def test_misra1a_reference_unit_tests(nist_problems):
    # unit_tests_gslnls.R:108-115: lm+trace, dogleg+levenberg, lmaccel+marquardt, lm+unit weights
    pr = nist_problems["Misra1a"]
    data = {k: np.array(v) for k, v in pr["data"].items()}
    rows = O.sympy_rows(O.split_formula(pr["formula"])[1], pr["param_names"], {"x": data["x"]})
    tol = np.finfo(float).eps ** 0.25
    cases = [dict(algorithm="lm", trace=True), dict(algorithm="dogleg", scale="levenberg"),
             dict(algorithm="lmaccel", scale="marquardt"), dict(algorithm="lm", weights=np.ones(14))]
    base = None
    for kw in cases:
        r = O.nls_large(rows, data["y"], pr["start"], **kw)
        assert np.max(np.abs(r["par"] - np.array(pr["target"]))) <= tol
        if kw.get("algorithm") == "lm":
            base = r if base is None else base
            assert np.array_equal(r["par"], base["par"])  # unit weights change nothing


def test_linear_full_rank():
    # unit_tests_gslnls.R:122-131 / src/test_nls.f90:390-482: f_i = x_i - 2 sum(x)/m - 1, m = n = 5
    n = 5
    Jc = np.eye(n) - 2.0 / n

    def rows(th, v, wf, wJ, wh):
        return th - 2.0 * np.sum(th) / n - 1.0, Jc, np.zeros(n)
    for alg in ("lm", "subspace2D", "cgst", "lmaccel"):
        r = O.nls_large(rows, np.zeros(n), np.ones(n), algorithm=alg)
        assert r["conv"] == 0
        assert np.max(np.abs(r["par"] + 1.0)) <= np.finfo(float).eps ** 0.25, alg


def test_weights_are_weighted_least_squares(readme_examples):
    # consistent sqrt(w) row scaling (src/fdf.c:153-167 semantics): integer weights == replication
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    w = 1.0 + (np.arange(x.size) % 3)
    r1 = O.nls_large("exp3", y, [1, 1, 0], x=x, weights=w)
    xr, yr = np.repeat(x, w.astype(int)), np.repeat(y, w.astype(int))
    r2 = O.nls_large("exp3", yr, [1, 1, 0], x=xr)
    assert np.allclose(r1["par"], r2["par"], rtol=1e-9)
    assert r1["ssr"] == pytest.approx(r2["ssr"], rel=1e-10)


def test_failure_conventions(readme_examples):
    # src/nls_large.c:298-302,319-326: non-success/non-maxiter status returns start and NA covar
    def rows(th, v, wf, wJ, wh):
        return np.full(4, np.nan), np.full((4, 2), np.nan), None
    r = O.nls_large(rows, np.zeros(4), [1.0, 2.0], algorithm="lm")
    assert r["conv"] == 9 and r["status"] == "problem with user-supplied function"
    assert np.array_equal(r["par"], [1.0, 2.0]) and np.all(np.isnan(r["covar"]))


def test_longdouble_accumulation_agrees(readme_examples):
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    a = O.eval_packet("exp3", y, [1.0, 1.0, 0.0], x=x)
    b = O.eval_packet("exp3", y, [1.0, 1.0, 0.0], x=x, longdouble=True)
    assert np.allclose(a, b, rtol=1e-14)
    assert a.size == 10


def test_thurber_iteration_count_is_not_pinned_by_the_algorithm(nist_problems):
    """Why tests/test_gpu_parity.py::test_nist_fits holds Thurber (NIST "higher difficulty", 7 parameters, start 1) to
    niter +- 3 and 5e-7 on the coefficients instead of equality and 1e-8: the oracle itself, with nothing changed
    but the order in which the 37 observations are added (rows permuted) or the width of the accumulators (long
    double), moves by up to three iterations and ~5e-8 in the coefficients, while SSR agrees to 1e-12 -- the
    trust-region path through this problem's flat valley is decided by roundings.  A CUDA kernel with yet another
    (fixed) summation order cannot be asked to reproduce one particular of these paths."""
    pr = nist_problems["Thurber"]
    x, y = np.array(pr["data"]["x"]), np.array(pr["data"]["y"])
    rhs = O.split_formula(pr["formula"])[1]
    perm = np.random.default_rng(1).permutation(y.size)
    spread_iter, spread_par = 0, 0.0
    for alg in ("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"):
        runs = []
        for xx, yy, ld in ((x, y, False), (x, y, True), (x[perm], y[perm], False)):
            r = O.nls_large(O.sympy_rows(rhs, pr["param_names"], {"x": xx}), yy, pr["start"], algorithm=alg, longdouble=ld)
            assert r["conv"] == 0
            runs.append(r)
        base = runs[0]
        for r in runs[1:]:
            assert abs(r["ssr"] - base["ssr"]) <= 1e-12 * base["ssr"]
            spread_iter = max(spread_iter, abs(r["niter"] - base["niter"]))
            spread_par = max(spread_par, float(np.max(np.abs(r["par"] - base["par"]) / np.abs(base["par"]))))
    assert 1 <= spread_iter <= 3          # observed: up to 3 (subspace2D 43 / 42 / 40)
    assert 1e-8 < spread_par < 5e-7       # observed: 5.4e-8
