"""GPU parity of the tiled pass kernel (K1b, p > 8: shared-memory J tiles + FP64 DMMA SYRK), through the
C ABI, against the CPU oracle.  BASELINE.json configs[3] shape: sum of K Gaussians, parameters (a, m, s)
per component; the oracle's row evaluator is orc_rows_gaussmix (oracle/dense_model.c)."""
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


def gaussmix_formula(K):
    terms, names = [], []
    for k in range(1, K + 1):
        terms.append("a%d * exp(-(x - m%d)^2 / s%d^2)" % (k, k, k))
        names += ["a%d" % k, "m%d" % k, "s%d" % k]
    return " + ".join(terms), names


def gaussmix_truth(K):
    th = []
    for k in range(1, K + 1):
        th += [5.0 + ((7 * k) % 11), 100.0 * (k - 0.5) / K, 2.5]
    return np.array(th)


def gaussmix_data(n, K, seed=3):
    """SURVEY 8(d) config 4 design: x on [0, 100], noise sd 0.5"""
    rng = np.random.Generator(np.random.Philox(key=seed))
    x = 100.0 * np.arange(n) / max(n - 1, 1)
    th = gaussmix_truth(K)
    y = np.zeros(n)
    for k in range(K):
        y += th[3 * k] * np.exp(-((x - th[3 * k + 1]) ** 2) / th[3 * k + 2] ** 2)
    return x, y + 0.5 * rng.standard_normal(n)


def rel_packet_err(got, ref, p):
    npk = p * (p + 1) // 2
    out = []
    for sl in (slice(0, npk), slice(npk, npk + p), slice(npk + p, npk + p + 1)):
        out.append(np.max(np.abs(got[sl] - ref[sl])) / np.max(np.abs(ref[sl])))
    return max(out)


@pytest.mark.parametrize("K", [3, 4, 5, 8, 16])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 100_003])
def test_packet_parity_gaussmix(G, K, n):
    p = 3 * K
    x, y = gaussmix_data(n, K)
    rhs, names = gaussmix_formula(K)
    m = G.Model(rhs, names, ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], y)
    theta = gaussmix_truth(K) * (1.0 + 0.02 * (-1.0) ** np.arange(p))
    got = pb.eval_packet(theta)
    ref = O.eval_packet("gaussmix", y, theta, x=x, longdouble=True)
    assert rel_packet_err(got, ref, p) < 1e-12, (K, n)
    again = pb.eval_packet(theta)
    assert np.array_equal(got, again)  # fixed-order reduction: bit identical
    pb.close()


def test_packet_parity_gaussmix_weights(G):
    K, n = 16, 20_011
    x, y = gaussmix_data(n, K)
    w = 0.5 + (np.arange(n) % 5) / 2.0
    rhs, names = gaussmix_formula(K)
    m = G.Model(rhs, names, ["x"], jac=True)
    pb = G.Problem(m, n, has_weights=True).upload([x], y, w)
    theta = gaussmix_truth(K) * 1.01
    got = pb.eval_packet(theta)
    ref = O.eval_packet("gaussmix", y, theta, x=x, weights=w, longdouble=True)
    assert rel_packet_err(got, ref, 3 * K) < 1e-12
    pb.close()


def test_jtfvv_and_fd_jacobian_tiled(G):
    """FVV mode (J^T fvv) and the finite-difference Jacobian variant of the tiled kernel, p = 12"""
    K, n = 4, 5003
    x, y = gaussmix_data(n, K)
    rhs, names = gaussmix_formula(K)
    theta = gaussmix_truth(K) * 1.01
    rng = np.random.default_rng(2)
    v = rng.standard_normal(3 * K)
    f = np.empty(n); J = np.empty((n, 3 * K)); h = np.empty(n)
    rows = O.sympy_rows(rhs, names, {"x": x})
    _, J, h = rows(theta, v, False, True, True)
    m = G.Model(rhs, names, ["x"], jac=True, fvv=True)
    pb = G.Problem(m, n).upload([x], y)
    ref = J.T @ h
    assert np.allclose(pb.eval_jtfvv(theta, v), ref, rtol=1e-9, atol=1e-9 * np.max(np.abs(ref)))
    pb.close()
    mfd = G.Model(rhs, names, ["x"], jac="forward")
    pb = G.Problem(mfd, n).upload([x], y)
    got = pb.eval_packet(theta)
    ref = O.eval_packet("gaussmix", y, theta, x=x, longdouble=True, fd_jac=1)
    assert rel_packet_err(got, ref, 3 * K) < 1e-9
    pb.close()


@pytest.mark.parametrize("alg", ["lm", "dogleg", "ddogleg", "lmaccel", "subspace2D", "cgst"])
def test_gaussmix48_fit_matches_oracle(G, alg):
    """config 4 at reduced n: p = 48, start = truth (1 +- 2 %)"""
    K, n = 16, 20_000
    p = 3 * K
    x, y = gaussmix_data(n, K)
    rhs, names = gaussmix_formula(K)
    start = gaussmix_truth(K) * (1.0 + 0.02 * (-1.0) ** np.arange(p))
    m = G.Model(rhs, names, ["x"], jac=True, fvv=(alg == "lmaccel"))
    pb = G.Problem(m, n).upload([x], y)
    fit = pb.fit(start, algorithm=alg)
    ref = O.nls_large("gaussmix", y, start, x=x, algorithm=alg)
    assert fit["conv"] == ref["conv"], (fit["status"], ref["status"])
    assert fit["niter"] == ref["niter"]
    assert np.allclose(fit["par"], ref["par"], rtol=1e-8)
    assert fit["ssr"] == pytest.approx(ref["ssr"], rel=1e-8)
    pb.close()


def test_tiled_exp_accuracy(G):
    """the tiled kernels evaluate exp() with the table-based branch-free nls_exp() (nls_model_prelude.h):
    K4 materialises f = exp(c x) for a p = 9 model whose other terms vanish; compare with numpy's exp"""
    n = 200_001
    x = np.concatenate([np.linspace(-745.0, 709.0, n - 7), [0.0, -0.0, 1e-300, -1e-300, 800.0, -800.0, 708.5]])
    names = ["c"] + ["z%d" % k for k in range(8)]
    rhs = "exp(c * x)" + "".join(" + z%d * x" % k for k in range(8))
    m = G.Model(rhs, names, ["x"], jac=True)
    pb = G.Problem(m, n).upload([x], np.zeros(n))
    f = pb.residuals([1.0] + [0.0] * 8)
    with np.errstate(over="ignore"):
        ref = np.exp(x)
    ok = (np.abs(x) < 708.0)
    rel = np.abs(f[ok] - ref[ok]) / ref[ok]
    assert rel.max() < 4 * np.finfo(float).eps, rel.max()   # <= 1.24 ulp by construction
    assert np.all(f[x >= 708.0] == np.inf) and np.all(f[x <= -708.0] == 0.0)
    pb.close()
