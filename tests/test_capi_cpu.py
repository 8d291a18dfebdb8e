"""CPU-side checks of the product library: it loads, exports every symbol the header declares,
compiles models with NVRTC for sm_100a (needs no GPU) and refuses to compute without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from gslnls_b200 import Model, _lib, gsl_nls_control, gsl_nls_large, pack_control

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gslnls_b200.h")).read()
    declared = set(re.findall(r"GSLNLS_API\s+[^;(]*?\b(gslnls_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r"\b(gslnls_\w+)\b", out))
    assert declared <= exported, declared - exported
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    assert b"sm_100a" in L.gslnls_version()
    assert L.gslnls_strerror(27) == b"iteration is not making progress towards solution"
    assert L.gslnls_strerror(11) == b"exceeded max number of iterations"
    assert L.gslnls_trs_name(1) == b"levenberg-marquardt+accel"  # README.md:649


def test_no_cuda_symbols_leak_and_no_torch_types():
    hdr = open(os.path.join(ROOT, "include", "gslnls_b200.h")).read()
    assert "torch" not in hdr and "at::" not in hdr


def test_control_defaults_and_packing():
    c = gsl_nls_control()
    assert len(c) == 23  # R/nls.R:1156 "exactly twenty-three components"
    assert c["maxiter"] == 100 and c["scale"] == "more" and c["solver"] == "qr" and c["fdtype"] == "forward"
    assert c["factor_up"] == 2 and c["factor_down"] == 3 and c["avmax"] == 0.75 and c["h_fvv"] == 0.02
    assert c["xtol"] == c["ftol"] == c["gtol"] == c["h_df"] == np.sqrt(np.finfo(float).eps)
    assert c["mstart_n"] == 30 and c["mstart_q"] == 3 and c["mstart_p"] == 5
    ci, cd = pack_control(c, "ddogleg", True)
    assert ci.tolist() == [100, 1, 3, 0, 0, -2, 0]          # R/nls_large.R:383-391
    assert cd.tolist() == [2, 3, 0.75, c["h_df"], 0.02, c["xtol"], c["ftol"], c["gtol"]]  # :407
    with pytest.raises(ValueError):
        gsl_nls_control(scale="nope")
    with pytest.raises(ValueError):
        gsl_nls_control(maxiter=0)
    with pytest.raises(ValueError):
        gsl_nls_control(xtol=-1)


def test_argument_validation_messages():
    x = np.linspace(0, 1, 8)
    y = 2 * np.exp(-x)
    kw = dict(data={"x": x, "y": y}, start={"A": 1.0, "lam": 1.0})
    with pytest.raises(ValueError, match="analytic Jacobian function 'jac' is required"):
        gsl_nls_large("y ~ A * exp(-lam * x)", **kw)
    with pytest.raises(ValueError, match="analytic second derivative function 'fvv' is required"):
        gsl_nls_large("y ~ A * exp(-lam * x)", algorithm="lmaccel", jac=True, **kw)
    with pytest.raises(ValueError, match="should be one of"):
        gsl_nls_large("y ~ A * exp(-lam * x)", algorithm="newton", jac=True, **kw)
    with pytest.raises(ValueError, match="negative residual degrees of freedom"):
        gsl_nls_large("y ~ A * exp(-lam * x)", data={"x": x[:1], "y": y[:1]}, start={"A": 1, "lam": 1}, jac=True)
    with pytest.raises(ValueError, match="parameters without starting value"):
        gsl_nls_large("y ~ A * exp(-lam * x) + b", jac=True, **kw)   # unit_tests_gslnls.R:118
    with pytest.raises(ValueError, match="starting values"):
        gsl_nls_large("y ~ A * exp(-lam * x)", data={"x": x, "y": y}, jac=True)  # :117
    with pytest.raises(ValueError, match="non-positive weights"):
        gsl_nls_large("y ~ A * exp(-lam * x)", jac=True, weights=np.zeros(8), **kw)
    with pytest.raises(TypeError):
        gsl_nls_large("y ~ A * exp(-lam * x)", data=[1, 2], start={"A": 1}, jac=True)


def test_untranslatable_models_fail_loudly():
    with pytest.raises(_lib.GslnlsError) as ei:
        Model("A * besselJ(x, 0)", ["A"], ["x"])
    assert ei.value.code == 1001
    with pytest.raises(_lib.GslnlsError):
        Model("A * exp(-lam * x", ["A", "lam"], ["x"])
    with pytest.raises(_lib.GslnlsError):
        Model("A * exp(-lam * z)", ["A", "lam"], ["x"])


@pytest.mark.skipif(_lib.lib().gslnls_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"])
    x = np.linspace(0, 3, 25)
    with pytest.raises(_lib.GslnlsError) as ei:
        gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": x}, start={"A": 1, "lam": 1, "b": 0},
                      jac=True, model=m)
    assert ei.value.code == 1004  # GSLNLS_ENODEVICE


def test_nvrtc_compiles_every_reference_formula_for_sm100a(nist_problems):
    """all 33 formula problems of R/nls_test.R translate, differentiate and compile (symbolic J + fvv)"""
    for name, pr in nist_problems.items():
        rhs = pr["formula"].split("~", 1)[1]
        lhs = pr["formula"].split("~", 1)[0]
        vars_ = [k for k in pr["data"] if k not in lhs.replace("log(", "").replace(")", "").split()]
        m = Model(rhs, pr["param_names"], vars_, jac=True, fvv=True)
        assert "nls_model_fvv" in m.source, name


@pytest.mark.parametrize("jac,fvv", [("forward", "fd"), ("center", None), (True, "fd")])
def test_nvrtc_compiles_finite_difference_modes(jac, fvv):
    m = Model("a * exp(-(x - b)^2 / (2 * c^2))", ["a", "b", "c"], ["x"], jac=jac, fvv=fvv)
    assert ("nls_model_fj" in m.source) == (jac is True)


def test_selfstart_shapes_translate():
    # inst/unit_tests/unit_tests_gslnls.R:272-275 uses SSasymp with gsl_nls_large
    m = Model("SSasymp(x, Asym, R0, lrc)", ["Asym", "R0", "lrc"], ["x"], jac=True, fvv=True)
    assert "NLS_EXP(" in m.source


@pytest.mark.skipif(_lib.lib().gslnls_device_count() > 0, reason="a GPU is present")
def test_every_fit_entry_point_refuses_without_a_device():
    """gslnls_fit_large / _sharded / _multi and the device-group constructor: a library error, no crash,
    no silent CPU path"""
    L = _lib.lib()
    m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True)
    x = np.linspace(0, 3, 32)
    y = 2 * np.exp(-x) + 1
    ci, cd = pack_control(gsl_nls_control(), "lm", False)
    st = np.array([1.0, 1.0, 0.0])
    arr = (_lib.c_double_p * 1)(x.ctypes.data_as(_lib.c_double_p))
    dp = _lib.c_double_p
    common = (y.size, st.ctypes.data_as(dp), ci.ctypes.data_as(_lib.c_int_p), cd.ctypes.data_as(dp))
    res = _lib.Result()
    rc = L.gslnls_fit_large(m.handle, arr, y.ctypes.data_as(dp), None, *common, 0, 0, C.byref(res))
    assert rc == 1004 and not res.par
    rc = L.gslnls_fit_large_sharded(m.handle, arr, y.ctypes.data_as(dp), None, *common, 0, None, 0, C.byref(res))
    assert rc == 1004
    dev = np.array([0, 1], dtype=np.int32)
    rc = L.gslnls_fit_large_multi(m.handle, arr, y.ctypes.data_as(dp), None, *common, 2,
                                  dev.ctypes.data_as(_lib.c_int_p), 0, C.byref(res))
    assert rc >= 1000
    out = (C.c_void_p * 2)()
    assert L.gslnls_comm_create_local(2, dev.ctypes.data_as(_lib.c_int_p), out) >= 1000
    assert not out[0] and not out[1]
    # fewer observations than parameters is an argument error before any device is touched (R/nls_large.R:286)
    rc = L.gslnls_fit_large_multi(m.handle, arr, y.ctypes.data_as(dp), None, 2, st.ctypes.data_as(dp),
                                  ci.ctypes.data_as(_lib.c_int_p), cd.ctypes.data_as(dp), 2,
                                  dev.ctypes.data_as(_lib.c_int_p), 0, C.byref(res))
    assert rc == 4 and b"degrees of freedom" in L.gslnls_last_error()
    L.gslnls_cache_clear()  # nothing cached: must be a no-op


K1_TUNES = [
    "tiled=0,block=256,unroll=3,minb=2,prefetch=1,fexp=1",      # LDG-pipelined default
    "tiled=0,block=256,unroll=4,minb=2,prefetch=0,fexp=0",      # plain LDG, library exp
    "tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1",        # TMA ring default
    "tiled=2,block=96,unroll=1,minb=4,stages=2,fexp=2",         # smallest ring, polynomial exp
]


@pytest.mark.parametrize("tune", K1_TUNES)
def test_nvrtc_compiles_every_k1_load_path_for_sm100a(monkeypatch, tune):
    """both load paths of the p <= 4 pass kernel (and the exp flavours) compile for sm_100a without a GPU;
    the developer override is how the GPU tests force each of them"""
    monkeypatch.setenv("GSLNLS_TUNE", tune)
    m = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    assert "nls_model_fj" in m.source
    m2 = Model("A * exp(-lam * x) + b", ["A", "lam", "b"], ["x"], jac="center", fvv="fd")
    assert "nls_model_f" in m2.source


def test_compiled_modules_are_cached_per_process(monkeypatch):
    """the host language compiles 'the same formula' on every .Call: the second NVRTC compile of an identical
    module (generated source + options) is served from the process-wide cache; a different module is not"""
    import time
    rhs = "A * exp(-lam * x) + b + 0.125"
    t0 = time.perf_counter()
    Model(rhs, ["A", "lam", "b"], ["x"], jac=True)
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    m2 = Model(rhs, ["A", "lam", "b"], ["x"], jac=True)
    second = time.perf_counter() - t0
    assert second < 0.25 * first and "nls_model_fj" in m2.source
    monkeypatch.setenv("GSLNLS_TUNE", "tiled=0,block=128,unroll=2,minb=2,prefetch=0,fexp=0")  # other options: a miss
    t0 = time.perf_counter()
    Model(rhs, ["A", "lam", "b"], ["x"], jac=True)
    assert time.perf_counter() - t0 > 4 * second
