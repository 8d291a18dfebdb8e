"""The drop-in boundary is a C ABI: a plain C99 program (no C++, no torch, no Python) must be able to include
include/gslnls_b200.h, link libgslnls_b200.so and run the one-call fit.  Without a GPU the call must refuse
(no CPU path); on a B200 it must give the oracle's answer for README Example 1."""
import os
import subprocess

import numpy as np
import pytest

from gslnls_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_consumer", "fit_example.c")


def _build(tmp_path):
    exe = str(tmp_path / "fit_example")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1",
                           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir, "-lgslnls_b200",
                           "-lm", "-Wl,-rpath," + libdir])
    return exe


def _write_data(tmp_path, x, y):
    path = tmp_path / "data.txt"
    with open(path, "w") as fh:
        fh.write("%d\n" % len(x))
        for a, b in zip(x, y):
            fh.write("%.17g %.17g\n" % (a, b))
    return str(path)


def _run(exe, data, ngpu=1):
    out = subprocess.run([exe, data, str(ngpu)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    kv = {}
    for ln in out.stdout.splitlines():
        k, _, v = ln.partition(" ")
        kv[k] = v
    return kv


@pytest.mark.skipif(_lib.lib().gslnls_device_count() > 0, reason="a GPU is present")
def test_c_program_compiles_links_and_is_refused_without_a_device(tmp_path, readme_examples):
    e = readme_examples["example1"]
    kv = _run(_build(tmp_path), _write_data(tmp_path, e["x"], e["y"]))
    assert kv["rc"] == "1004" and "no usable CUDA device" in kv["error"]


@pytest.mark.gpu
def test_c_program_fits_example1(tmp_path, readme_examples):
    from oracle import oracle as O
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"])
    kv = _run(_build(tmp_path), _write_data(tmp_path, x, y))
    ref = O.nls_large("exp3", y, [0.0, 0.0, 0.0], x=x, algorithm="lm")
    assert int(kv["rc"]) == ref["conv"] == 0 and int(kv["niter"]) == ref["niter"]
    assert kv["status"] == "success" and kv["algorithm"] == "levenberg-marquardt"
    par = [float(v) for v in kv["par"].split()]
    assert np.allclose(par, ref["par"], rtol=1e-8)
    assert float(kv["ssr"]) == pytest.approx(ref["ssr"], rel=1e-8)
    assert float(kv["resid_ss"]) == pytest.approx(float(kv["ssr"]), rel=1e-12)
    # README.md:187-194: A = 4.893, lam = 1.417, b = 1.010
    assert np.allclose(par, [4.893, 1.417, 1.010], atol=5e-4)
