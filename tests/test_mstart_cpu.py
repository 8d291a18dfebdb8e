"""Multi-start global search without a GPU: the product's control logic (gslnls_b200/csrc/mstart.hpp, compiled
for the host by tests/host_harness over the host build of the trust-region core) against the oracle's
line-by-line restatement of gsl_multistart_driver (oracle/mstart.py), on the fixtures of the reference's own
tests (inst/unit_tests/unit_tests_gslnls.R:137-176: BoxBOD and Madsen with mstart_n = 5, mstart_q = 1,
mstart_r = 1.1), plus the quasi-random generators against their published sequence heads."""
import numpy as np
import pytest

import trs_host as T
from oracle import mstart as OM
from oracle import oracle as O


def test_sobol_matches_published_sequence_and_oracle():
    # Bratley & Fox / gsl_qrng_sobol: the sequence starts at (0.5, ..), never at the origin; first dimension is
    # the van der Corput sequence in Gray-code order; leading points of dimensions 2 and 3 as published
    pts = T.qrng(3, 8)
    assert np.array_equal(pts[:4], [[0.5, 0.5, 0.5], [0.75, 0.25, 0.75], [0.25, 0.75, 0.25], [0.375, 0.375, 0.625]])
    assert np.array_equal(pts[4:8, 0], [0.875, 0.625, 0.125, 0.1875])
    for dim in (1, 2, 4, 7, 13, 20):
        a = T.qrng(dim, 300)
        g = OM.Sobol(dim)
        b = np.array([g.next() for _ in range(300)])
        assert np.array_equal(a, b), dim                      # two independent restatements, bit for bit
        # a (t, m, s)-net property every Sobol' sequence has: each coordinate of the first 2^k points is a
        # permutation of the odd multiples of 2^-(k+1) ... together with the skipped origin, all multiples of 2^-k
        k = 8
        for d in range(dim):
            assert sorted(np.concatenate([[0.0], a[:2 ** k - 1, d]]) * 2 ** k) == list(range(2 ** k)), (dim, d)
    # scipy's generator uses Joe-Kuo direction numbers: dimensions 1 and 2 coincide with Bratley-Fox
    from scipy.stats import qmc
    ref = qmc.Sobol(2, scramble=False).random(257)[1:]
    assert np.array_equal(T.qrng(2, 256), ref)


def test_halton_is_the_radical_inverse():
    pts = T.qrng(41, 5)   # more than 40 parameters: gsl_qrng_halton (src/nls.c:279-280)
    assert np.allclose(pts[:3, 0], [0.5, 0.25, 0.75])
    assert np.allclose(pts[:3, 1], [1 / 3, 2 / 3, 1 / 9])
    assert np.allclose(pts[0, 40], 1.0 / 179)                # 41st prime


def _boxbod(nist_problems):
    pr = nist_problems["BoxBOD"]
    data = {k: np.array(v) for k, v in pr["data"].items()}
    rows = O.sympy_rows(O.split_formula(pr["formula"])[1], pr["param_names"], {"x": data["x"]})
    return pr, data, rows


def _madsen():
    # inst/unit_tests/unit_tests_gslnls.R:66-74 (Madsen et al. example): 3 residuals, 2 parameters
    y = np.zeros(3)

    def rows(theta, v, wf, wJ, wh):
        x1, x2 = theta
        f = np.array([x1 ** 2 + x2 ** 2 + x1 * x2, np.sin(x1), np.cos(x2)])
        J = np.array([[2 * x1 + x2, 2 * x2 + x1], [np.cos(x1), 0.0], [0.0, -np.sin(x2)]])
        return f, J, np.zeros(3)
    return y, rows, np.array([-0.155437, 0.694564])


CASES = {
    # name: (ranges, has_range)
    "boxbod_4.1.1": ([[200.0, 250.0], [0.0, 1.0]], [[1, 1], [1, 1]]),
    "boxbod_point_and_range": ([[200.0, 250.0], [1.0, 1.0]], [[1, 1], [1, 1]]),          # 4.1.3
    "boxbod_unknown_b1": ([[-0.1, 0.75], [0.0, 1.0]], [[0, 0], [1, 1]]),                  # 4.1.6 without bounds
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_boxbod_control_logic_matches_oracle_and_finds_the_certified_minimum(nist_problems, case):
    pr, data, rows = _boxbod(nist_problems)
    rng, has = CASES[case]
    kw = dict(mstart_n=5, mstart_q=1, mstart_r=1.1 * (10 if not np.all(has) else 1))     # R/nls.R:711-713
    prov = T.packet_from_rows(rows, data["y"])
    got = T.multistart(prov, rng, has, n=data["y"].size, **kw)
    ref = OM.multistart(rows, data["y"], rng, has, **kw)
    assert got["status"] == ref["status"] and got["mstarts"] == ref["mstarts"], (got, ref)
    assert got["nsp"] == ref["nsp"] and got["nwsp"] == ref["nwsp"]
    assert np.allclose(got["par"], ref["par"], rtol=1e-6), (got["par"], ref["par"])
    assert got["ssr"] == pytest.approx(ref["ssr"], rel=1e-6)
    assert np.allclose(got["range"], ref["range"], rtol=1e-9)
    # the final fit from the multi-start optimum lands on the certified values (the reference's own check)
    fit = O.nls_large(rows, data["y"], got["par"], algorithm="lm")
    assert fit["conv"] == 0
    assert np.max(np.abs(fit["par"] - np.array(pr["target"])) / np.abs(pr["target"])) < 1e-6


@pytest.mark.parametrize("rng,has", [([[-1.0, 1.0], [0.0, 1.0]], [[1, 1], [1, 1]]),      # 4.2.1
                                     ([[-0.1, 0.75], [-0.1, 0.75]], [[0, 0], [0, 0]])])   # 4.2.4: no ranges at all
def test_madsen_control_logic_matches_oracle(rng, has):
    y, rows, target = _madsen()
    kw = dict(mstart_n=5, mstart_q=1, mstart_r=1.1 * (10 if not np.all(has) else 1))
    prov = T.packet_from_rows(rows, y)
    got = T.multistart(prov, rng, has, n=3, **kw)
    ref = OM.multistart(rows, y, rng, has, **kw)
    assert got["status"] == ref["status"] and got["mstarts"] == ref["mstarts"], (got, ref)
    assert got["nsp"] == ref["nsp"]
    assert np.allclose(got["par"], ref["par"], rtol=1e-6, atol=1e-9)
    fit = O.nls_large(rows, y, got["par"], algorithm="lm")
    assert np.allclose(fit["par"], target, atol=np.finfo(float).eps ** 0.25)   # unit_tests dotest_tol
