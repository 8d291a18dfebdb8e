"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs.  Tolerances are BASELINE.json's: per-evaluation J^T J / J^T f
within 1e-12 relative, final coefficients / SSR / iteration count within 1e-8 relative."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
import os  # noqa: E402

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALGS = ["lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"]


@pytest.fixture(scope="module")
def G():
    import gslnls_b200
    from gslnls_b200 import _lib
    assert _lib.lib().gslnls_device_count() > 0, "no CUDA device: the product path has no fallback"
    return gslnls_b200


def synth_exp(n, seed=1):
    rng = np.random.Generator(np.random.Philox(key=seed))
    x = 3.0 * np.arange(n) / max(n - 1, 1)
    y = 5.0 * np.exp(-1.5 * x) + 1.0 + 0.25 * rng.standard_normal(n)
    return x, y


def rel_packet_err(got, ref, p):
    """1e-12 gate: relative to the magnitude of each block (JTJ, JTf, fTf)"""
    npk = p * (p + 1) // 2
    out = []
    for sl in (slice(0, npk), slice(npk, npk + p), slice(npk + p, npk + p + 1)):
        out.append(np.max(np.abs(got[sl] - ref[sl])) / np.max(np.abs(ref[sl])))
    return max(out)


# ---------------------------------------------------------------- K1: packet parity
@pytest.mark.parametrize("n", [1, 2, 3, 7, 25, 255, 256, 257, 1000, 4097, 65536, 1_000_003])
def test_packet_parity_exp3(G, n):
    x, y = synth_exp(n)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, n).upload([x], y)
    for theta in ([1.0, 1.0, 0.0], [5.0, 1.5, 1.0], [0.0, 0.0, 0.0]):
        got = pb.eval_packet(theta)
        ref = O.eval_packet("exp3", y, theta, x=x, longdouble=True)
        assert rel_packet_err(got, ref, 3) < 1e-12, (n, theta)
    pb.close()


def test_packet_parity_weights_and_unaligned(G):
    n = 10001
    x, y = synth_exp(n)
    w = 0.5 + (np.arange(n) % 7) / 3.0
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    pb = G.Problem(m, n, has_weights=True).upload([x], y, w)
    theta = [4.0, 1.2, 0.8]
    got = pb.eval_packet(theta)
    ref = O.eval_packet("exp3", y, theta, x=x, weights=w, longdouble=True)
    assert rel_packet_err(got, ref, 3) < 1e-12
    pb.close()
    # 8-byte aligned device columns (a shard starting at an odd row) take the scalar-load variant
    import torch
    xt = torch.tensor(np.concatenate([[0.0], x]), device="cuda")
    yt = torch.tensor(np.concatenate([[0.0], y]), device="cuda")
    pb = G.Problem(m, n).bind_device([xt.data_ptr() + 8], yt.data_ptr() + 8, keepalive=[xt, yt])
    got = pb.eval_packet(theta)
    ref = O.eval_packet("exp3", y, theta, x=x, longdouble=True)
    assert rel_packet_err(got, ref, 3) < 1e-12
    pb.close()


def test_packet_deterministic_and_nonfinite_rule(G):
    n = 300_001
    x, y = synth_exp(n)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], y)
    a = pb.eval_packet([2.0, 0.7, 0.1])
    for _ in range(3):
        assert np.array_equal(a, pb.eval_packet([2.0, 0.7, 0.1]))  # fixed-order reduction: bit identical
    # model value overflows to +Inf -> residual +Inf (src/nls_large.c:464-465) -> f^T f = +Inf
    bad = pb.eval_packet([1.0, -1e6, 0.0])
    assert np.isinf(bad[-1]) and bad[-1] > 0
    pb.close()


def test_packet_parity_nist_models(G, nist_problems):
    rng = np.random.default_rng(5)
    for name in ["Thurber", "Gauss3", "Misra1a", "Lubricant", "Nelson", "ENSO", "Hahn1", "MGH09"]:
        pr = nist_problems[name]
        lhs, rhs = O.split_formula(pr["formula"])
        vars_ = [k for k in pr["data"] if k not in lhs.replace("log(", "").replace(")", "").split()]
        data = {k: np.array(pr["data"][k]) for k in vars_}
        y = np.log(np.array(pr["data"]["y"])) if lhs.startswith("log") else np.array(pr["data"]["y"])
        m = G.Model(rhs, pr["param_names"], vars_, jac=True, fvv=True)
        pb = G.Problem(m, y.size).upload([data[k] for k in vars_], y)
        theta = np.array(pr["target"]) * (1 + 0.01 * rng.standard_normal(pr["p"]))
        got = pb.eval_packet(theta)
        rows = O.sympy_rows(rhs, pr["param_names"], data)
        ref = O.eval_packet(rows, y, theta, longdouble=True)
        assert rel_packet_err(got, ref, pr["p"]) < 5e-12, name
        v = rng.standard_normal(pr["p"])
        _, J, h = rows(theta, v, False, True, True)
        assert np.allclose(pb.eval_jtfvv(theta, v), J.T @ h, rtol=1e-9, atol=1e-9 * np.max(np.abs(J.T @ h))), name
        pb.close()


# ---------------------------------------------------------------- full fits
def _fit_cmp(fit, ref, tol=1e-8):
    assert fit["conv"] == ref["conv"], (fit["status"], ref["status"])
    assert fit["niter"] == ref["niter"]
    assert np.allclose(fit["par"], ref["par"], rtol=tol, atol=0)
    assert fit["ssr"] == pytest.approx(ref["ssr"], rel=tol)


@pytest.mark.parametrize("alg", ALGS)
def test_example1_all_methods(G, readme_examples, alg):
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    start = [1.0, 1.0, 0.0]
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, x.size).upload([x], y)
    fit = pb.fit(start, algorithm=alg, trace=True, want_resid_grad=True)
    ref = O.nls_large("exp3", y, start, x=x, algorithm=alg, trace=True, want_resid_grad=True)
    _fit_cmp(fit, ref)
    k = ref["niter"] + 1
    assert np.allclose(fit["ssrtrace"][:k], ref["ssrtrace"], rtol=1e-8)
    assert np.allclose(fit["partrace"][:k], ref["partrace"], rtol=1e-7, atol=1e-10)
    assert np.allclose(fit["covar"], ref["covar"], rtol=1e-7)
    # evaluated at parameters that agree to 1e-8
    assert np.allclose(fit["resid"], ref["resid"], rtol=1e-6, atol=1e-7)
    assert np.allclose(fit["grad"], ref["grad"], rtol=1e-6, atol=1e-7)
    # the final trial steps sit at rounding level (||f_trial|| >= ||f|| by an ulp or not): the last
    # iteration may end by acceptance on one side and by 16 rejections on the other, same niter
    assert abs(fit["neval"]["df2"] - ref["neval"]["df2"]) <= 1
    assert abs(fit["neval"]["fvv"] - ref["neval"]["fvv"]) <= 17
    pb.close()


def test_example1_readme_start_and_object(G, readme_examples):
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    fit = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start={"A": 0, "lam": 0, "b": 0},
                          jac=True, trace=True)
    assert fit.convInfo["isConv"] and fit.convInfo["finIter"] == e["niter"]           # README.md:193
    assert [round(v, 3) for v in fit.coef().values()] == e["coef_print"]             # README.md:187-189
    assert round(fit.deviance(), 3) == e["ssr_print"]
    se = [round(s["std_error"], 4) for s in fit.summary()["coefficients"].values()]
    assert se == e["stderr_print"]                                                    # README.md:252-254
    assert round(fit.sigma(), 4) == e["sigma_print"]
    assert fit.convInfo["trsName"] == "multilarge/levenberg-marquardt"
    assert np.allclose(fit.fitted() + fit.residuals(), y)
    assert np.allclose(fit.predict({"x": x}), fit.fitted())
    assert fit.partrace.shape == (e["niter"] + 1, 3)


def test_example2_readme_traces_on_gpu(G, readme_examples):
    """README.md:772-804: lmaccel with analytic fvv, 12 iterations, every printed digit"""
    e = readme_examples["example2"]
    x, y = np.array(e["x"]), np.array(e["y"])
    fit = G.gsl_nls_large("y ~ a * exp(-(x - b)^2 / (2 * c^2))", data={"x": x, "y": y},
                          start={"a": 1, "b": 0, "c": 1}, algorithm="lmaccel", jac="forward", fvv=True, trace=True)
    g = e["lmaccel_fvv"]
    assert fit.convInfo["finIter"] == g["niter"]
    for t in g["trace"]:
        assert fit.devtrace[t["iter"]] == pytest.approx(t["ssr"], rel=6e-6)
        assert np.allclose(fit.partrace[t["iter"]], t["par"], rtol=6e-6, atol=1e-6)
    # forward-difference Jacobian + FD fvv (README.md:636-659)
    fit = G.gsl_nls_large("y ~ a * exp(-(x - b)^2 / (2 * c^2))", data={"x": x, "y": y},
                          start={"a": 1, "b": 0, "c": 1}, algorithm="lmaccel", jac="forward", fvv="fd", trace=True)
    g = e["lmaccel_fd"]
    assert fit.convInfo["finIter"] == g["niter"]
    for t in g["trace"]:
        assert fit.devtrace[t["iter"]] == pytest.approx(t["ssr"], rel=6e-6)
    # lm with forward differences: 26 iterations (README.md:568-616)
    fit = G.gsl_nls_large("y ~ a * exp(-(x - b)^2 / (2 * c^2))", data={"x": x, "y": y},
                          start={"a": 1, "b": 0, "c": 1}, algorithm="lm", jac="forward", trace=True)
    assert fit.convInfo["finIter"] == e["lm_fd"]["niter"]
    assert [round(v, 4) for v in fit.coef().values()] == e["coef_print"]


@pytest.mark.parametrize("name", ["Misra1a", "Thurber", "Gauss3", "Chwirut2", "Kirby2", "BoxBOD"])
def test_nist_fits(G, nist_problems, name):
    """NIST StRD via gsl_nls_large, checked against certified values (config 2) and the oracle"""
    pr = nist_problems[name]
    data = {k: np.array(v) for k, v in pr["data"].items()}
    rows = O.sympy_rows(O.split_formula(pr["formula"])[1], pr["param_names"], {"x": data["x"]})
    for alg in ALGS:
        ref = O.nls_large(rows, data["y"], pr["start"], algorithm=alg)
        fit = G.gsl_nls_large(pr["formula"], data=data, start=dict(zip(pr["param_names"], pr["start"])),
                              algorithm=alg, jac=True, fvv=True if alg == "lmaccel" else None)
        got = np.array(list(fit.coef().values()))
        assert fit.convInfo["stopCode"] == ref["conv"], (name, alg)
        if ref["conv"] == 0 and name != "BoxBOD":
            # Thurber: the oracle against itself (rows permuted / long-double accumulators) differs by up to 3
            # iterations and 5.4e-8 in the coefficients -- tests/test_oracle_pins.py::
            # test_thurber_iteration_count_is_not_pinned_by_the_algorithm; the gate is 10 x that sensitivity
            hard = name in ("Thurber",)
            assert abs(fit.convInfo["finIter"] - ref["niter"]) <= (3 if hard else 0), (name, alg)
            assert np.allclose(got, ref["par"], rtol=5e-7 if hard else 1e-8), (name, alg)
            assert fit.deviance() == pytest.approx(ref["ssr"], rel=1e-8)
            assert np.max(np.abs(got - np.array(pr["target"])) / np.abs(pr["target"])) < 1e-6, (name, alg)


def test_reference_unit_tests_3_1(G, nist_problems):
    """inst/unit_tests/unit_tests_gslnls.R:108-115 (tolerance .Machine$double.eps^0.25)"""
    pr = nist_problems["Misra1a"]
    data = {k: np.array(v) for k, v in pr["data"].items()}
    st = dict(zip(pr["param_names"], pr["start"]))
    tol = np.finfo(float).eps ** 0.25
    cases = [dict(jac=True, trace=True),
             dict(algorithm="dogleg", jac=True, control={"scale": "levenberg"}),
             dict(algorithm="lmaccel", jac=True, fvv=True, control={"scale": "marquardt"}),
             dict(algorithm="lm", weights=np.ones(14), jac=True)]
    for kw in cases:
        fit = G.gsl_nls_large(pr["formula"], data=data, start=st, **kw)
        assert np.max(np.abs(np.array(list(fit.coef().values())) - np.array(pr["target"]))) <= tol, kw


def test_weighted_fit_matches_oracle(G, readme_examples):
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    w = 1.0 + (np.arange(x.size) % 3)
    fit = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start={"A": 1, "lam": 1, "b": 0},
                          jac=True, weights=w)
    ref = O.nls_large("exp3", y, [1, 1, 0], x=x, weights=w)
    assert fit.convInfo["finIter"] == ref["niter"]
    assert np.allclose(list(fit.coef().values()), ref["par"], rtol=1e-8)
    assert fit.deviance() == pytest.approx(ref["ssr"], rel=1e-8)


def test_failure_returns_start_and_na(G):
    x = np.linspace(0, 1, 16)
    y = np.ones(16)
    m = G.Model("A * log(-x - lam)", ["A", "lam"], ["x"], jac=True)  # log of a negative number: NaN Jacobian
    pb = G.Problem(m, 16).upload([x], y)
    fit = pb.fit([1.0, 2.0], want_resid_grad=True)
    assert fit["conv"] == 9 and fit["status"] == "problem with user-supplied function"
    assert np.array_equal(fit["par"], [1.0, 2.0]) and np.all(np.isnan(fit["covar"]))
    assert np.all(np.isnan(fit["resid"])) and np.all(np.isnan(fit["grad"]))      # src/nls_large.c:345-376
    pb.close()


# ---------------------------------------------------------------- size-independent properties at scale
def test_large_n_properties(G):
    n = 20_000_000
    x, y = synth_exp(n)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, n).upload([x], y)
    th = np.array([4.0, 1.3, 0.9])
    full = pb.eval_packet(th)
    # additivity over row blocks (the property the multi-GPU sharding relies on)
    parts = np.zeros_like(full)
    for lo, hi in ((0, 7_000_001), (7_000_001, 13_000_000), (13_000_000, n)):
        q = G.Problem(m, hi - lo).upload([x[lo:hi]], y[lo:hi])
        parts += q.eval_packet(th)
        q.close()
    assert rel_packet_err(parts, full, 3) < 1e-12
    # J^T J entries that have closed forms: sum 1 = n, sum J0 = sum exp(-lam x)
    assert full[5] == n
    assert full[3] == pytest.approx(np.sum(np.exp(-th[1] * x)), rel=1e-12)
    # f^T f against numpy pairwise summation
    r = th[0] * np.exp(-th[1] * x) + th[2] - y
    assert full[9] == pytest.approx(np.sum(r * r), rel=1e-12)
    fit = pb.fit([1.0, 1.0, 0.0], algorithm="lm")
    assert fit["conv"] == 0
    assert np.allclose(fit["par"], [5.0, 1.5, 1.0], rtol=2e-3)
    assert np.max(np.abs(fit["grad_vec"])) < 1e-3 * fit["ssr"]
    ref = O.nls_large("exp3", y, [1.0, 1.0, 0.0], x=x, algorithm="lm", threads=8)
    _fit_cmp(fit, ref)
    pb.close()


def test_batched_multistart_inner_loops(G):
    """config 5 shape at test size: many start points, 5 LM iterations each + log det(J^T J) screen"""
    rng = np.random.Generator(np.random.Philox(key=2))
    n, S = 1024, 257
    x = np.linspace(0, 10, n)
    y = 3 * np.exp(-0.5 * x) + 2 * np.exp(-3 * x) + 0.05 * rng.standard_normal(n)
    starts = rng.uniform(0.1, 5.0, (S, 4))
    m = G.Model("A1*exp(-l1*x)+A2*exp(-l2*x)", ["A1", "l1", "A2", "l2"], ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], y)
    out = pb.fit_batch(starts, iters=5)
    for c in range(0, S, 16):
        ref = O.nls_large("expmix2", y, starts[c], x=x, algorithm="lm", maxiter=5)
        if ref["conv"] in (0, 11):
            assert np.allclose(out["par"][c], ref["par"], rtol=1e-6, atol=1e-9), c
            assert out["ssr"][c] == pytest.approx(ref["ssr"], rel=1e-7)
        pk = O.eval_packet("expmix2", y, starts[c], x=x)
        JTJ = np.zeros((4, 4))
        JTJ[np.tril_indices(4)] = pk[:10]
        JTJ = JTJ + np.tril(JTJ, -1).T
        sign, ld = np.linalg.slogdet(JTJ)
        if sign > 0 and np.isfinite(out["logdet"][c]) and np.linalg.cond(JTJ) < 1e8:
            assert out["logdet"][c] == pytest.approx(ld, rel=1e-6, abs=1e-6)
    pb.close()


# ---------------------------------------------------------------- K1 load-path variants
# The library picks the load path of the p <= 4 kernel by shard size: software-pipelined LDG below
# ~1 GB per pass, the TMA bulk-copy shared-memory ring above.  Both must give the oracle's packet at
# every size, ragged tails included, so each is forced here through the developer override.
K1_VARIANTS = {
    "ldg-prefetch": "tiled=0,block=256,unroll=3,minb=2,prefetch=1,fexp=1",
    "ldg-plain-libexp": "tiled=0,block=256,unroll=4,minb=2,prefetch=0,fexp=0",
    "tma-ring": "tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1",
    "tma-ring-small": "tiled=2,block=96,unroll=1,minb=4,stages=2,fexp=2",
}


@pytest.mark.parametrize("variant", sorted(K1_VARIANTS))
@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 2303, 2304, 2305, 4608, 100_003, 1_000_003])
def test_packet_parity_k1_variants(G, monkeypatch, variant, n):
    monkeypatch.setenv("GSLNLS_TUNE", K1_VARIANTS[variant])
    x, y = synth_exp(n)
    w = 0.5 + (np.arange(n) % 7) / 3.0
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    for weights in (None, w):
        pb = G.Problem(m, n, has_weights=weights is not None).upload([x], y, weights)
        theta = [4.0, 1.2, 0.8]
        got = pb.eval_packet(theta)
        ref = O.eval_packet("exp3", y, theta, x=x, weights=weights, longdouble=True)
        assert rel_packet_err(got, ref, 3) < 1e-12, (variant, n, weights is not None)
        assert np.array_equal(got, pb.eval_packet(theta))  # run-to-run bitwise
        pb.close()


@pytest.mark.parametrize("variant", ["tma-ring", "ldg-prefetch"])
@pytest.mark.parametrize("alg", ["lm", "lmaccel", "cgst"])
def test_fit_k1_variants(G, monkeypatch, variant, alg):
    """full fits through each load path (FJ, FVV and JVP pass modes) against the oracle"""
    monkeypatch.setenv("GSLNLS_TUNE", K1_VARIANTS[variant])
    n = 300_007
    x, y = synth_exp(n)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, n).upload([x], y)
    fit = pb.fit([1.0, 1.0, 0.0], algorithm=alg)
    ref = O.nls_large("exp3", y, [1.0, 1.0, 0.0], x=x, algorithm=alg)
    _fit_cmp(fit, ref)
    pb.close()


def test_table_exp_special_values(G):
    """the branch-free exp of the default kernels on overflow / underflow / NaN arguments: the
    non-finite rule of src/nls_large.c:464-465 must see exactly what the library exp() would show it"""
    n = 64
    x = np.linspace(-1.0, 1.0, n)
    y = np.zeros(n)
    m = G.Model("A * exp(lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], y)
    for theta in ([1.0, 800.0, 0.0], [1.0, 5000.0, 0.0], [1.0, 1e308, 0.0]):
        got = pb.eval_packet(theta)
        with np.errstate(over="ignore"):
            f = theta[0] * np.exp(theta[1] * x) + theta[2]
        nbad = int(np.sum(~np.isfinite(f)))
        # packet layout: [JTJ(6) | JTf(3) | fTf]; any non-finite residual makes fTf = +Inf
        assert (not np.isfinite(got[9])) == (nbad > 0), (theta, got[9], nbad)
    pb.close()


# ---------------------------------------------------------------- one-shot call and its cache
def _one_shot(G, m, x, y, start, alg="lm", weights=None, want_rg=False):
    import ctypes as C
    from gslnls_b200 import _lib
    ci, cd = G.pack_control(G.gsl_nls_control(), alg, False)
    st = np.ascontiguousarray(start, dtype=np.float64)
    arr = (_lib.c_double_p * 1)(x.ctypes.data_as(_lib.c_double_p))
    res = _lib.Result()
    rc = _lib.lib().gslnls_fit_large(m.handle, arr, y.ctypes.data_as(_lib.c_double_p),
                                     weights.ctypes.data_as(_lib.c_double_p) if weights is not None else None,
                                     y.size, st.ctypes.data_as(_lib.c_double_p), ci.ctypes.data_as(_lib.c_int_p),
                                     cd.ctypes.data_as(_lib.c_double_p), 0, int(want_rg), C.byref(res))
    _lib.check(rc)
    out = {"par": [res.par[i] for i in range(st.size)], "ssr": res.ssr, "niter": res.niter, "conv": res.conv,
           "n": res.n, "status": res.status.decode()}
    if want_rg:
        out["resid"] = np.ctypeslib.as_array(res.resid, shape=(y.size,)).copy()
    _lib.lib().gslnls_result_free(C.byref(res))
    return out


def test_one_shot_call_reuses_cached_buffers_correctly(G):
    """gslnls_fit_large() keeps device buffers / workspace / streams per device between calls: results must
    not depend on what the previous call was (smaller n, larger n, weights on and off, another model)"""
    from gslnls_b200 import _lib
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    m2 = G.Model("A * exp(-lam * x)", ["A", "lam"], ["x"], jac=True)
    start = [1.0, 1.0, 0.0]
    seen = {}
    for rnd in range(2):
        for n in (50_001, 1_000, 200_000, 25):
            x, y = synth_exp(n, seed=n)
            w = 0.5 + (np.arange(n) % 5) / 4.0
            for weights in (None, w):
                got = _one_shot(G, m, x, y, start, weights=weights, want_rg=(n <= 1000))
                key = (n, weights is not None)
                if key not in seen:
                    ref = O.nls_large("exp3", y, start, x=x, algorithm="lm", weights=weights)
                    assert got["conv"] == ref["conv"] and got["niter"] == ref["niter"], key
                    assert np.allclose(got["par"], ref["par"], rtol=1e-8), key
                    assert got["ssr"] == pytest.approx(ref["ssr"], rel=1e-8)
                    if "resid" in got:
                        assert abs(np.sum(got["resid"] ** 2) - got["ssr"]) <= 1e-10 * got["ssr"]
                    seen[key] = got
                else:  # second round: bitwise the same answer out of recycled buffers
                    assert got["par"] == seen[key]["par"] and got["ssr"] == seen[key]["ssr"], key
            other = _one_shot(G, m2, x, y - 1.0, [1.0, 1.0])  # another model evicts the cached problem
            assert other["n"] == n
    _lib.lib().gslnls_cache_clear()
    x, y = synth_exp(1000, seed=1000)
    again = _one_shot(G, m, x, y, start)
    assert again["par"] == seen[(1000, False)]["par"]
    # freeing a model drops what was cached for it (no dangling kernels): a fresh model fits again
    del m2
    m3 = G.Model("A * exp(-lam * x)", ["A", "lam"], ["x"], jac=True)
    assert _one_shot(G, m3, x, y - 1.0, [1.0, 1.0])["conv"] in (0, 11, 27)


def test_tma_ring_shrinks_to_fit_many_columns(G, monkeypatch):
    """three predictors + response + weights = 5 columns: the default 4-stage ring of 2304-row tiles would
    need 369 KB of shared memory; the library must shrink it (or fall back) and still match the oracle"""
    monkeypatch.setenv("GSLNLS_TUNE", K1_VARIANTS["tma-ring"])
    rng = np.random.Generator(np.random.Philox(key=11))
    n = 200_003
    x1, x2, x3 = rng.uniform(0, 2, n), rng.uniform(-1, 1, n), rng.uniform(0.5, 1.5, n)
    y = 2.0 * np.exp(-0.7 * x1) + 0.5 * x2 * x3 + 0.1 * rng.standard_normal(n)
    w = 0.5 + (np.arange(n) % 3) / 2.0
    m = G.Model("a * exp(-b * x1) + c * x2 * x3", ["a", "b", "c"], ["x1", "x2", "x3"], jac=True)
    pb = G.Problem(m, n, has_weights=True).upload([x1, x2, x3], y, w)
    a, b, c = theta = np.array([1.5, 0.5, 0.3])
    got = pb.eval_packet(theta)
    e = np.exp(-b * x1)
    sw = np.sqrt(w)
    r = (a * e + c * x2 * x3 - y) * sw
    Jw = np.stack([e, -a * x1 * e, x2 * x3], axis=1) * sw[:, None]
    JTJ = Jw.T @ Jw
    ref = np.concatenate([JTJ[np.tril_indices(3)], Jw.T @ r, [r @ r]])
    assert rel_packet_err(got, ref, 3) < 1e-10  # numpy double sums here, not the long-double oracle
    fit = pb.fit([1.0, 1.0, 0.0])
    assert fit["conv"] == 0 and np.allclose(fit["par"], [2.0, 0.7, 0.5], rtol=2e-2)
    pb.close()


# ---------------------------------------------------------------- round-2 additions
def test_center_difference_jacobian_on_gpu(G, readme_examples):
    """jac="center": src/fdjac.c:81-128 (+-delta/2, delta = h |x_j| or h) inside the pass kernel, against the
    oracle's restatement of the same rule.  A difference quotient with step delta ~ 1.5e-8 amplifies the last-ulp
    differences between the device exp() and libm's by 1/delta: J entries agree to ~eps/h ~ 1e-8 by
    construction, so the gates here are 5e-8 on the packet (not 1e-12, which is a summation gate) and 1e-6 on
    the fitted parameters, with equal iteration counts."""
    e = readme_examples["example2"]
    x, y = np.array(e["x"]), np.array(e["y"])
    m = G.Model("a * exp(-(x - b)^2 / (2 * c^2))", ["a", "b", "c"], ["x"], jac="center")
    pb = G.Problem(m, x.size).upload([x], y)
    for theta in ([1.0, 0.0, 1.0], [4.5, 0.45, 0.15]):
        got = pb.eval_packet(theta)
        ref = O.eval_packet("gauss", y, theta, x=x, fd_jac=2, longdouble=True)
        assert rel_packet_err(got, ref, 3) < 5e-8, theta
        ana = O.eval_packet("gauss", y, theta, x=x, longdouble=True)
        assert rel_packet_err(got, ana, 3) < 1e-6   # and close to the analytic Jacobian's packet
    for alg in ("lm", "dogleg"):
        fit = pb.fit(e["start"], algorithm=alg, control=dict(G.gsl_nls_control(), fdtype="center"))
        ref = O.nls_large("gauss", y, e["start"], x=x, algorithm=alg, fd_jac=2, fdtype="center")
        _fit_cmp(fit, ref, tol=1e-6)
    pb.close()
    # tiled kernel (p > 8) with centred differences
    rng = np.random.Generator(np.random.Philox(key=21))
    n, K = 5000, 3
    xx = np.linspace(0, 30, n)
    th = np.array([5.0, 5.0, 2.0, 7.0, 15.0, 2.5, 4.0, 25.0, 2.0])
    yy = sum(th[3 * k] * np.exp(-((xx - th[3 * k + 1]) ** 2) / th[3 * k + 2] ** 2) for k in range(K))
    yy = yy + 0.1 * rng.standard_normal(n)
    rhs = " + ".join("a%d * exp(-(x - m%d)^2 / s%d^2)" % (k, k, k) for k in range(1, K + 1))
    names = [s % k for k in range(1, K + 1) for s in ("a%d", "m%d", "s%d")]
    m9 = G.Model(rhs, names, ["x"], jac="center")
    pb = G.Problem(m9, n).upload([xx], yy)
    st = th * 1.02
    got = pb.eval_packet(st)
    ref = O.eval_packet("gaussmix", yy, st, x=xx, fd_jac=2, longdouble=True)
    assert rel_packet_err(got, ref, 9) < 5e-7
    pb.close()


@pytest.mark.parametrize("alg", ["lm", "dogleg"])
def test_condtrace_matches_oracle_on_gpu(G, nist_problems, alg):
    """cond(J) column of the trace (callback_large, src/nls_large.c:733-738): GSL's cholesky_rcond estimator on
    the device against the oracle's restatement, p = 3 (resident server) and p = 7 / 8 (cooperative solves)"""
    for name in ("Misra1a", "Thurber", "Gauss3"):
        pr = nist_problems[name]
        data = {k: np.array(v) for k, v in pr["data"].items()}
        rows = O.sympy_rows(O.split_formula(pr["formula"])[1], pr["param_names"], {"x": data["x"]})
        ref = O.nls_large(rows, data["y"], pr["start"], algorithm=alg, maxiter=8, trace=True)
        m = G.Model(O.split_formula(pr["formula"])[1], pr["param_names"], ["x"], jac=True)
        pb = G.Problem(m, data["y"].size).upload([data["x"]], data["y"])
        fit = pb.fit(pr["start"], algorithm=alg, control=dict(G.gsl_nls_control(), maxiter=8), trace=True)
        k = min(fit["niter"], ref["niter"]) + 1
        assert k >= 3, name
        assert np.allclose(fit["condtrace"][1:k], ref["condtrace"][1:k], rtol=1e-5), (name, alg)
        pb.close()


def test_weights_mode_gsl_matches_reference_semantics(G, readme_examples):
    """weights_mode="gsl": sqrt(w) on f only, J^T J unweighted -- what the reference computes for non-unit
    weights (libgsl's winit never sees J; gsl_df_large, src/nls_large.c:474-653, does not weight it)"""
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    w = 1.0 + (np.arange(x.size) % 3)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, x.size, has_weights=True).set_weights_mode("gsl").upload([x], y, w)
    theta = [4.0, 1.2, 0.8]
    got = pb.eval_packet(theta)
    ref = O.eval_packet("exp3", y, theta, x=x, weights=w, longdouble=True, weights_gsl=True)
    assert rel_packet_err(got, ref, 3) < 1e-12
    other = O.eval_packet("exp3", y, theta, x=x, weights=w, longdouble=True)
    assert rel_packet_err(got, other, 3) > 1e-3   # the two semantics really differ
    for alg in ("lm", "lmaccel", "dogleg"):
        fit = pb.fit([1.0, 1.0, 0.0], algorithm=alg, want_resid_grad=True)
        ref = O.nls_large("exp3", y, [1.0, 1.0, 0.0], x=x, weights=w, algorithm=alg, weights_gsl=True,
                          want_resid_grad=True)
        _fit_cmp(fit, ref)
        assert np.allclose(fit["covar"], ref["covar"], rtol=1e-7)
        assert np.allclose(fit["resid"], ref["resid"], rtol=1e-6, atol=1e-9)   # weighted residuals
        assert np.allclose(fit["grad"], ref["grad"], rtol=1e-6, atol=1e-9)     # unweighted Jacobian (:354-363)
    pb.close()
    hi = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start={"A": 1, "lam": 1, "b": 0},
                         jac=True, weights=w, weights_mode="gsl")
    ref = O.nls_large("exp3", y, [1, 1, 0], x=x, weights=w, weights_gsl=True)
    assert hi.convInfo["finIter"] == ref["niter"] and np.allclose(list(hi.coef().values()), ref["par"], rtol=1e-8)


def test_inf_residual_finite_jacobian_on_gpu(G):
    """model value overflows while the Jacobian stays finite: the reference's NaN scan covers J only
    (src/nls_large.c:515-522), so the fit iterates on with residual +Inf (16 rejected trials -> ENOPROG in the
    first iteration, par = start) instead of stopping with EBADFUNC"""
    x = np.linspace(0.0, 1.0, 64)
    y = 1.0 + 2.0 * x

    def rows(theta, v, wf, wJ, wh):
        A, b = theta
        with np.errstate(all="ignore"):
            f = A * x + b + 1e308 * (1.0 + np.sign(A - 5.0))
        return f, np.stack([x, np.ones_like(x)], axis=1), np.zeros_like(x)
    m = G.Model("A * x + b + 1e308 * (1 + sign(A - 5))", ["A", "b"], ["x"], jac=True, fvv=True)
    pb = G.Problem(m, x.size).upload([x], y)
    for alg in ALGS:
        for st in ([10.0, 0.0], [4.9, 0.0]):
            fit = pb.fit(st, algorithm=alg)
            ref = O.nls_large(rows, y, st, algorithm=alg)
            assert fit["conv"] == ref["conv"] == (27 if st[0] > 5 else 0), (alg, st, fit["status"])
            assert fit["niter"] == ref["niter"], (alg, st)
            assert np.allclose(fit["par"], ref["par"], rtol=1e-8), (alg, st)
    pb.close()


def test_serialised_execution_falls_back_to_launch_ordered(tmp_path):
    """CUDA_LAUNCH_BLOCKING=1 (and ncu / compute-sanitizer) serialise kernels: the resident trust-region
    server cannot run next to the pass kernel.  The library must notice -- by environment, or through the
    server's start-of-fit handshake when told to try anyway (GSLNLS_SERVER=1) -- and step launch-ordered."""
    import json
    import subprocess
    import sys
    code = r'''
import json, sys, time
import numpy as np
sys.path.insert(0, %r)
import gslnls_b200 as G
rng = np.random.Generator(np.random.Philox(key=1))
n = 200001
x = 3.0 * np.arange(n) / (n - 1)
y = 5.0 * np.exp(-1.5 * x) + 1.0 + 0.25 * rng.standard_normal(n)
m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
pb = G.Problem(m, n).upload([x], y)
out = {}
for alg in ("lm", "lmaccel"):
    t0 = time.time()
    f = pb.fit([1.0, 1.0, 0.0], algorithm=alg)
    out[alg] = {"par": list(f["par"]), "niter": f["niter"], "conv": f["conv"], "ssr": f["ssr"], "s": time.time() - t0}
print(json.dumps(out))
''' % ROOT_DIR
    results = {}
    for tag, env_extra in (("plain", {}), ("blocking", {"CUDA_LAUNCH_BLOCKING": "1"}),
                           ("blocking_forced_server", {"CUDA_LAUNCH_BLOCKING": "1", "GSLNLS_SERVER": "1",
                                                       "GSLNLS_WATCHDOG_S": "20"})):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, (tag, r.stderr[-2000:])
        results[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    for tag in ("blocking", "blocking_forced_server"):
        for alg in ("lm", "lmaccel"):
            a, b = results[tag][alg], results["plain"][alg]
            assert a["conv"] == b["conv"] == 0 and a["niter"] == b["niter"], (tag, alg)
            assert np.allclose(a["par"], b["par"], rtol=1e-12) and a["ssr"] == pytest.approx(b["ssr"], rel=1e-12)
            assert a["s"] < 10.0, (tag, alg, a["s"])   # a handshake period, not a 60 s watchdog


def test_upload_from_pageable_memory_is_exact_and_fast(G):
    """gslnls_problem_upload stages pageable host arrays through the library's pinned ring (upload.cpp): the
    device copy must be bit-exact at sizes that are not multiples of the slice, and repeatable"""
    import time
    n = 6_000_011
    x, y = synth_exp(n)
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    pb = G.Problem(m, n)
    t0 = time.perf_counter()
    pb.upload([x], y)
    r = pb.residuals([0.0, 0.0, 1.0])   # A = 0: residual = 1 - y, i.e. the uploaded y column itself
    dt = time.perf_counter() - t0
    assert np.array_equal(r, 1.0 - y)
    r2 = pb.residuals([1.0, 0.0, 0.0])  # lam = 0: residual = 1 - y as well; x enters through J only
    assert np.array_equal(r2, 1.0 - y)
    _, J = pb.residuals([1.0, 1.0, 0.0], want_grad=True)
    assert np.allclose(J[:, 0], np.exp(-x), rtol=1e-14)   # the uploaded x column, through the model
    assert dt < 5.0
    pb.close()


# ---------------------------------------------------------------- multi-start
def test_multistart_boxbod_madsen_on_gpu(G, nist_problems):
    """gslnls_problem_multistart: control logic of gsl_multistart_driver over the batched kernels, against the
    oracle's restatement on the reference's fixtures (inst/unit_tests/unit_tests_gslnls.R:137-176)"""
    from oracle import mstart as OM
    pr = nist_problems["BoxBOD"]
    data = {k: np.array(v) for k, v in pr["data"].items()}
    rhs = O.split_formula(pr["formula"])[1]
    rows = O.sympy_rows(rhs, pr["param_names"], {"x": data["x"]})
    m = G.Model(rhs, pr["param_names"], ["x"], jac=True)
    pb = G.Problem(m, data["y"].size).upload([data["x"]], data["y"])
    ctl = dict(mstart_n=5, mstart_q=1, mstart_r=1.1)
    for rng, has in (([[200.0, 250.0], [0.0, 1.0]], [[1, 1], [1, 1]]),
                     ([[-0.1, 0.75], [0.0, 1.0]], [[0, 0], [1, 1]])):
        got = pb.multistart(rng, has, control=ctl)
        ref = OM.multistart(rows, data["y"], rng, has, mstart_n=5, mstart_q=1,
                            mstart_r=1.1 * (10 if not np.all(has) else 1))
        assert got["status"] == ref["status"] and got["mstarts"] == ref["mstarts"], (got, ref)
        assert got["nsp"] == ref["nsp"] and np.allclose(got["par"], ref["par"], rtol=1e-6)
        fit = pb.fit(got["par"])
        assert fit["conv"] == 0 and np.max(np.abs(fit["par"] / np.array(pr["target"]) - 1)) < 1e-6
    pb.close()
    # the high-level call: ranges in `start` trigger the search, the final fit starts from its optimum
    obj = G.gsl_nls_large(pr["formula"], data=data, start={"b1": (200, 250), "b2": (0, 1)}, jac=True, control=ctl)
    assert obj.mstart["nsp"] >= 1 and obj.convInfo["isConv"]
    assert np.max(np.abs(np.array(list(obj.coef().values())) / np.array(pr["target"]) - 1)) < 1e-6
    obj = G.gsl_nls_large(pr["formula"], data=data, start={"b1": None, "b2": (0, 1)}, jac=True, control=ctl)
    assert np.max(np.abs(np.array(list(obj.coef().values())) / np.array(pr["target"]) - 1)) < 1e-6


def test_multistart_config5_shape_finds_the_global_minimum(G):
    """BASELINE configs[4] as a search: 8192 start points per major iteration batched per kernel, exponential
    mixture with known truth (3, 0.5, 2, 3); the search must end on the global minimiser (or its label swap)"""
    import bench
    x, y = bench.mstart_problem()
    m = G.Model("A1*exp(-l1*x)+A2*exp(-l2*x)", ["A1", "l1", "A2", "l2"], ["x"], jac=True)
    pb = G.Problem(m, x.size).upload([x], y)
    got = pb.multistart([[0.0, 10.0]] * 4, control=dict(mstart_n=8192, mstart_q=819, mstart_maxstart=8))
    fit = pb.fit(got["par"])
    assert fit["conv"] == 0
    par = fit["par"]
    if par[1] > par[3]:
        par = par[[2, 3, 0, 1]]
    assert np.allclose(par, [3.0, 0.5, 2.0, 3.0], rtol=0.05)
    assert fit["ssr"] < 1.02 * np.sum((3 * np.exp(-0.5 * x) + 2 * np.exp(-3 * x) - y) ** 2)
    assert got["searches"] >= 8192
    pb.close()


# ---------------------------------------------------------------- IRLS robust losses
def test_device_median_is_exact(G):
    """radix select over the bit patterns of |fn(theta) - y|: bitwise the median gsl_median() would return
    (src/nls_utils.c:162-189), odd and even lengths, ties, an infinite residual"""
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    theta = [4.0, 1.2, 0.8]
    for n in (1, 2, 3, 10, 255, 256, 1001, 65536, 300_007):
        x, y = synth_exp(n, seed=n)
        if n >= 10:
            y[: n // 3] = np.round(y[: n // 3], 1)        # ties
        pb = G.Problem(m, n).upload([x], y)
        r = np.abs(theta[0] * np.exp(-theta[1] * x) + theta[2] - y)
        got = pb.median_abs_resid(theta)
        s = np.sort(r)
        want = s[(n - 1) // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2.0
        assert got == pytest.approx(want, rel=1e-13), n    # the device exp differs from numpy's by an ulp
        pb.close()


@pytest.mark.parametrize("loss", ["huber", "barron", "bisquare", "welsh", "optimal", "hampel", "ggw", "lqq"])
def test_irls_matches_oracle(G, readme_examples, loss):
    from oracle import irls as OI
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"]).copy()
    y[[3, 11, 19]] += [6.0, -5.0, 8.0]
    rows = lambda th: th[0] * np.exp(-th[1] * x) + th[2]  # noqa: E731
    m = G.Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    for userw in (None, 1.0 + (np.arange(x.size) % 3)):
        pb = G.Problem(m, x.size, has_weights=True).upload([x], y, np.ones(x.size) if userw is None else userw)
        fit, info = pb.fit_irls([1.0, 1.0, 0.0], loss=loss)
        ref, rinfo = OI.irls("exp3", y, [1.0, 1.0, 0.0], loss=loss, x=x, rows=rows, weights=userw)
        assert info["status"] == rinfo["status"] == 0 and info["niter"] == rinfo["niter"], (loss, info, rinfo)
        # the outer iteration stops at irls_xtol = eps^(1/4) = 1.2e-4: both sides walk the same iterates (same
        # count), and each weighted fit agrees to its own 1e-8; the smooth losses carry that through at ~1e-7
        assert info["sigma"] == pytest.approx(rinfo["sigma"], rel=1e-6)
        assert np.allclose(fit["par"], ref["par"], rtol=1e-6), loss
        assert np.allclose(pb.weights(), rinfo["weights"], rtol=1e-6, atol=1e-12), loss
        pb.close()


def test_irls_high_level_and_scale(G):
    """gsl_nls_large(loss=...) on a contaminated large sample: the robust fit recovers the truth that the
    least-squares fit misses; sigma is the MAD scale of the clean noise"""
    n = 400_003
    x, y = synth_exp(n)
    rng = np.random.Generator(np.random.Philox(key=9))
    bad = rng.choice(n, n // 10, replace=False)
    y[bad] += 5.0 + 3.0 * rng.random(bad.size)             # 10 % one-sided outliers
    start = {"A": 1.0, "lam": 1.0, "b": 0.0}
    ls = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start=start, jac=True)
    rb = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start=start, jac=True, loss="bisquare")
    truth = np.array([5.0, 1.5, 1.0])
    err_ls = np.abs(np.array(list(ls.coef().values())) - truth)
    err_rb = np.abs(np.array(list(rb.coef().values())) - truth)
    assert rb.irls["status"] == 0 and rb.irls["niter"] >= 2
    assert err_rb[2] < 0.02 and err_ls[2] > 0.4             # the intercept absorbs the outliers in the LS fit
    # MAD scale of a sample with 10 % gross outliers: 1.4826 x the (0.5 / 0.9) quantile of |N(0, 0.25^2)|
    from scipy import stats
    assert rb.irls["sigma"] == pytest.approx(1.4826 * 0.25 * stats.norm.ppf(0.5 + 0.5 * 0.5 / 0.9), rel=0.02)
