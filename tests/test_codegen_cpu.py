"""The generated model functions are plain C++ as well as CUDA: compile them for the host and check
f, the symbolic Jacobian and the directional second derivative against sympy (an independent
differentiator) on every formula problem of the reference."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from gslnls_b200 import Model
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WRAP = r"""
extern "C" void eval_rows(const double* th, const double* v, const double* X, long n, double* f, double* J, double* h)
{
    for (long i = 0; i < n; ++i) {
        double x[GSLNLS_NVAR > 0 ? GSLNLS_NVAR : 1];
        for (int k = 0; k < GSLNLS_NVAR; ++k) x[k] = X[k * n + i];
        double fi, Ji[GSLNLS_P];
        nls_model_fj(th, x, fi, Ji);
        f[i] = fi;
        for (int j = 0; j < GSLNLS_P; ++j) J[i * GSLNLS_P + j] = Ji[j];
        h[i] = nls_model_fvv(th, v, x);
        if (nls_model_f(th, x) != fi && fi == fi) f[i] = 1e300; /* f and fj must agree bit for bit */
    }
}
"""


def host_eval(model, theta, v, cols):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "m.cpp")
        so = os.path.join(td, "m.so")
        open(src, "w").write(model.source + WRAP)
        subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", "-I",
                               os.path.join(ROOT, "gslnls_b200", "csrc"), src, "-o", so])
        L = C.CDLL(so)
        n = cols[0].size if cols else 1
        X = np.ascontiguousarray(np.stack(cols) if cols else np.zeros((1, n)))
        f, J, h = np.empty(n), np.empty((n, len(theta))), np.empty(n)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        th = np.ascontiguousarray(theta, dtype=float)
        vv = np.ascontiguousarray(v, dtype=float)
        L.eval_rows(dp(th), dp(vv), dp(X), C.c_long(n), dp(f), dp(J), dp(h))
        return f, J, h


def test_generated_code_matches_sympy_on_reference_formulas(nist_problems):
    rng = np.random.default_rng(7)
    checked = 0
    for name, pr in nist_problems.items():
        lhs, rhs = O.split_formula(pr["formula"])
        vars_ = [k for k in pr["data"] if k not in lhs.replace("log(", "").replace(")", "").split()]
        data = {k: np.array(pr["data"][k]) for k in vars_}
        theta = np.array(pr["target"]) * (1 + 0.01 * rng.standard_normal(pr["p"]))
        v = rng.standard_normal(pr["p"])
        m = Model(rhs, pr["param_names"], vars_, jac=True, fvv=True)
        f, J, h = host_eval(m, theta, v, [data[k] for k in vars_])
        rows = O.sympy_rows(rhs, pr["param_names"], data)
        f0, J0, h0 = rows(theta, v, True, True, True)
        assert np.allclose(f, f0, rtol=1e-12, atol=1e-13 * np.max(np.abs(f0))), name
        assert np.allclose(J, J0, rtol=1e-9, atol=1e-11 * np.max(np.abs(J0), axis=0)), name
        assert np.allclose(h, h0, rtol=1e-7, atol=1e-9 * np.max(np.abs(h0)) + 1e-300), name
        checked += 1
    assert checked == 33


@pytest.mark.parametrize("rhs,params,vars_", [
    ("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"]),
    ("a * exp(-(x - b)^2 / (2 * c^2))", ["a", "b", "c"], ["x"]),
    ("b1 * x**b2 + sin(b3 * x) / (1 + x^2) - log1p(b1^2) * tanh(b2 * x)", ["b1", "b2", "b3"], ["x"]),
    ("pnorm((x - m) / s) + dnorm(x, ) * 0 + atan(m * x) + sqrt(s) + x^-2 + 2^(m*x/10)", ["m", "s"], ["x"]),
    ("SSlogis(x, Asym, xmid, scal) + SSmicmen(z, Vm, K)", ["Asym", "xmid", "scal", "Vm", "K"], ["x", "z"]),
])
def test_generated_code_matches_sympy_misc(rhs, params, vars_):
    if "dnorm(x, )" in rhs:
        rhs = rhs.replace("dnorm(x, ) * 0 + ", "")
    rng = np.random.default_rng(3)
    data = {k: rng.uniform(0.5, 2.0, 40) for k in vars_}
    theta = rng.uniform(0.5, 1.5, len(params))
    v = rng.standard_normal(len(params))
    m = Model(rhs, params, vars_, jac=True, fvv=True)
    f, J, h = host_eval(m, theta, v, [data[k] for k in vars_])
    import sympy as sp
    srhs = rhs
    if "SSlogis" in rhs:
        srhs = "Asym/(1+exp((xmid-x)/scal)) + Vm*z/(K+z)"
    if "pnorm" in rhs:
        srhs = rhs.replace("pnorm(", "PN(").replace("log1p(", "LP(")
    loc = {"PN": lambda a: (1 + sp.erf(a / sp.sqrt(2))) / 2, "LP": lambda a: sp.log(1 + a)}
    names = params + vars_
    syms = {k: sp.Symbol(k, real=True) for k in names}
    loc.update(syms)
    e = sp.sympify(srhs.replace("^", "**").replace("log1p(", "LP("), locals=loc)
    ps = [syms[k] for k in params]
    args = list(theta) + [data[k] for k in vars_]
    lam = lambda ex: np.broadcast_to(sp.lambdify([syms[k] for k in names], ex, ["scipy", "numpy"])(*args), (40,))  # noqa: E731
    assert np.allclose(f, lam(e), rtol=1e-12)
    for j, s in enumerate(ps):
        assert np.allclose(J[:, j], lam(sp.diff(e, s)), rtol=1e-9, atol=1e-12), j
    hh = sum(v[i] * v[j] * sp.diff(e, ps[i], ps[j]) for i in range(len(ps)) for j in range(len(ps)))
    assert np.allclose(h, lam(hh), rtol=1e-8, atol=1e-10)
