"""r-package/src/*.c -- the .Call shim -- compiled against stub R headers (tests/r_stub) and driven the way
R/nls_large_cuda.R drives it: routines looked up in the registration table, SEXP arguments in, named list out,
residuals / gradient through the lazy accessor, explicit release + finalizer.  R itself is not installed in
the image; the stub implements the slice of R's C API the shim uses, so the shim cannot rot unnoticed."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")
DRIVER = os.path.join(STUB, "_build", "shim_driver")
NAMES = ["par", "covar", "resid", "grad", "niter", "status", "conv", "ssr", "ssrtol", "algorithm", "neval",
         "partrace", "ssrtrace", "jtj", "handle"]


def _build():
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.check_call(["make", "-C", STUB, "-s"], env=env)
    assert os.path.exists(DRIVER)


def _write(path, x, y, w=None):
    with open(path, "w") as fh:
        fh.write("%d %d\n" % (len(x), 0 if w is None else 1))
        for i in range(len(x)):
            fh.write("%.17g %.17g%s\n" % (x[i], y[i], "" if w is None else " %.17g" % w[i]))


def test_shim_is_small_compiles_warning_free_and_refuses_without_a_gpu(tmp_path, readme_examples):
    src = os.path.join(ROOT, "r-package", "src", "nls_large_cuda.c")
    assert sum(1 for _ in open(src)) < 200          # a shim, not a second implementation
    assert "oracle" not in open(src).read()
    _build()                                          # -Wall -Wextra -Werror against the stub headers
    import gslnls_b200._lib as L
    if L.lib().gslnls_device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test below")
    e = readme_examples["example1"]
    f = tmp_path / "ex1.txt"
    _write(f, e["x"], e["y"])
    r = subprocess.run([DRIVER, str(f), "0", "0"], capture_output=True, text=True, timeout=120)
    # the shim turns the library's GSLNLS_ENODEVICE into an R error (stub: exit status 3): no CPU fallback
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr, (r.returncode, r.stderr)


def test_r_front_end_packs_control_like_the_reference():
    """static checks of r-package/R/nls_large_cuda.R against R/nls_large.R:383-407 (R cannot run here)"""
    txt = open(os.path.join(ROOT, "r-package", "R", "nls_large_cuda.R")).read()
    for piece in ('c("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst")', "jacclass = -2L", "jacnz = 0L",
                  '"factor_up", "factor_down", "avmax", "h_df", "h_fvv", "xtol", "ftol", "gtol"',
                  'class(out) <- c("gsl_nls", "nls")', "C_nls_large_cuda_eval", 'paste("multilarge", cFit$algorithm'):
        assert piece in txt, piece
    ns = open(os.path.join(ROOT, "r-package", "NAMESPACE")).read()
    assert "useDynLib(gslnlscuda, .registration = TRUE)" in ns
    init = open(os.path.join(ROOT, "r-package", "src", "init.c")).read()
    assert '{"C_nls_large_cuda", (DL_FUNC)&C_nls_large_cuda, 11}' in init
    assert '{"C_nls_large_cuda_sparse", (DL_FUNC)&C_nls_large_cuda_sparse, 8}' in init
    sp = open(os.path.join(ROOT, "r-package", "R", "nls_large_cuda_sparse.R")).read()
    for piece in ("C_nls_large_cuda_sparse", '.cuda_pack_control(control, algorithm, trace)', "negative residual degrees of freedom",
                  'class(out) <- c("gsl_nls", "nls")'):
        assert piece in sp, piece
    assert "export(gsl_nls_large_cuda_sparse)" in ns and "export(nls_block)" in ns


@pytest.mark.gpu
@pytest.mark.parametrize("alg,wmode", [(0, 0), (1, 0), (2, 0), (0, 1)])
def test_shim_fits_example1_on_gpu(tmp_path, readme_examples, alg, wmode):
    from oracle import oracle as O
    _build()
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    w = 1.0 + (np.arange(x.size) % 3) if wmode else None
    f = tmp_path / "ex1.txt"
    _write(f, x, y, w)
    r = subprocess.run([DRIVER, str(f), str(alg), str(wmode)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    name = ["lm", "lmaccel", "dogleg"][alg]
    ref = O.nls_large("exp3", y, [1.0, 1.0, 0.0], x=x, algorithm=name, weights=w, weights_gsl=bool(wmode), trace=True,
                      want_resid_grad=True)
    assert d["names"] == NAMES                                     # src/nls_large.c:279-288 (+ jtj, handle)
    assert d["conv"] == ref["conv"] == 0 and d["niter"] == ref["niter"] and d["status"] == "success"
    assert np.allclose(d["par"], ref["par"], rtol=1e-8) and d["parnames"] == "lam"
    assert d["ssr"] == pytest.approx(ref["ssr"], rel=1e-8)
    assert np.allclose(np.array(d["covar"]).reshape(3, 3), ref["covar"], rtol=1e-7)
    assert d["resid_is_null"] == 1 and d["grad_is_null"] == 1     # lazy: the fit returns no O(n) arrays
    assert d["ntrace"] == 101 and d["partrace_dim"] == [101, 3] and d["ssrtrace0"] == pytest.approx(ref["ssrtrace"][0])
    assert d["resid_ss"] == pytest.approx(ref["ssr"], rel=1e-8)  # lazy residuals: sum of squares = ssr
    assert d["grad_dim"] == [x.size, 3]
    assert d["grad00"] == pytest.approx(ref["grad"][0, 0], rel=1e-7)
    assert d["grad_last"] == pytest.approx(ref["grad"][-1, 2], rel=1e-7)
    assert d["released"] == 1
    assert d["algorithm"] == ["levenberg-marquardt", "levenberg-marquardt+accel", "dogleg"][alg]


SPARSE_DRIVER = os.path.join(STUB, "_build", "sparse_driver")


def test_sparse_shim_is_small_and_refuses_without_a_gpu():
    """r-package/src/nls_large_cuda_sparse.c: the .Call shim of the sparse-Jacobian path"""
    src = os.path.join(ROOT, "r-package", "src", "nls_large_cuda_sparse.c")
    assert sum(1 for _ in open(src)) < 150 and "oracle" not in open(src).read()
    _build()
    import gslnls_b200._lib as L
    if L.lib().gslnls_device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test below")
    r = subprocess.run([SPARSE_DRIVER, "10", "0", "100"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("p,start,maxiter", [(10, 0, 100), (500, 1, 500)])
def test_sparse_shim_fits_penalty_on_gpu(p, start, maxiter, readme_examples):
    """inst/unit_tests/unit_tests_gslnls.R:316-346 (p = 10) and README Example 4 (p = 500) through the R shim"""
    import math
    from oracle import oracle as O
    _build()
    r = subprocess.run([SPARSE_DRIVER, str(p), str(start), str(maxiter)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["names"] == ["par", "niter", "status", "conv", "ssr", "ssrtol", "neval", "ssrtrace", "grad_vec", "jtj",
                            "resid", "cg_iters", "nnz"]
    assert out["conv"] == 0 and out["status"] == "success" and out["nnz"] == 2 * p
    assert out["ntrace"] == out["niter"] + 1 and abs(out["resid_ss"] - out["ssr"]) <= 1e-12 * out["ssr"]
    assert out["jtj_is_null"] == (p > 64)
    sa = math.sqrt(1e-5)
    eye = np.eye(p) * sa

    def rows(th, v, wf, wJ, wh):
        return (np.concatenate([sa * (th - 1), [np.sum(th ** 2)]]), np.vstack([eye, 2 * th[None, :]]) if wJ else None, None)
    y = np.zeros(p + 1)
    y[p] = 0.25
    st = np.arange(1, p + 1, dtype=float) if start else np.full(p, 0.15)
    ref = O.nls_large(rows, y, st, algorithm="cgst", maxiter=maxiter)
    assert out["niter"] == ref["niter"] and abs(out["ssr"] - ref["ssr"]) <= 1e-8 * ref["ssr"]
    if p == 500:
        assert float("%.7g" % out["ssr"]) == readme_examples["example4_penalty"]["ssr_print"]
    else:
        assert np.max(np.abs(np.array(out["par"]) - ref["par"]) / np.abs(ref["par"])) < 1e-8
