"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): observation shards + packet all-reduce must give
every rank the single-GPU / oracle answer."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from gslnls_b200 import _lib
    return _lib.lib().gslnls_device_count()


@pytest.mark.parametrize("world,p2p", [(2, 1), (2, 0), (4, 1), (8, 1)])
def test_sharded_fit_matches_oracle(tmp_path, world, p2p):
    """p2p=1: pass kernels deposit packets in every GPU's mailbox over NVLink peer memory and the resident
    trust-region warp sums them in rank order; p2p=0: NCCL all-gather + rank-order sum between kernels"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    import bench
    from oracle import oracle as O
    n = 2_000_003
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, GSLNLS_TEST_OUT=str(tmp_path), GSLNLS_TEST_N=str(n), GSLNLS_P2P=str(p2p),
               GSLNLS_WATCHDOG_S="20")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                           str(world), "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(ROOT, "tests", "dist_gpu_worker.py")], env=env, timeout=600)
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(world)]
    x, y = bench.synth_rows(0, n, n)
    full = O.eval_packet("exp3", y, [4.0, 1.3, 0.9], x=x, longdouble=True)
    for r in res:
        assert r["p2p"] == p2p
        assert r["packet"] == res[0]["packet"]            # all ranks: bitwise identical reduced packet
        assert r["fits"] == res[0]["fits"]                # ... and therefore identical trajectories
    got = np.array(res[0]["packet"])
    assert np.max(np.abs(got - full) / np.abs(full)) < 1e-12
    for alg, f in res[0]["fits"].items():
        ref = O.nls_large("exp3", y, list(bench.START), x=x, algorithm=alg)
        assert f["conv"] == ref["conv"] == 0 and f["niter"] == ref["niter"], alg
        assert np.allclose(f["par"], ref["par"], rtol=1e-8), alg
        assert f["ssr"] == pytest.approx(ref["ssr"], rel=1e-8)
        assert f["n"] == n


@pytest.mark.parametrize("ngpu", [2, 4, 8])
@pytest.mark.parametrize("alg", ["lm", "lmaccel", "dogleg"])
def test_single_process_multi_gpu_call(ngpu, alg):
    """gslnls_fit_large_multi(): the one-call form an R session would use -- host arrays in, the library
    splits the rows over the GPUs (one host thread per GPU, peer-memory mailboxes); same answer as the
    oracle and as the single-GPU call, residuals/gradient stitched back to n rows"""
    if _ngpu() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    import bench
    from gslnls_b200 import gsl_nls_large
    from oracle import oracle as O
    n = 1_000_003
    x, y = bench.synth_rows(0, n, n)
    kw = dict(data={"x": x, "y": y}, start=dict(zip(["A", "lam", "b"], bench.START)), jac=True, fvv=True,
              algorithm=alg)
    multi = gsl_nls_large("y ~ A * exp(-lam * x) + b", devices=list(range(ngpu)), **kw)
    single = gsl_nls_large("y ~ A * exp(-lam * x) + b", **kw)
    ref = O.nls_large("exp3", y, list(bench.START), x=x, algorithm=alg)
    assert multi.convInfo["isConv"] and multi.convInfo["finIter"] == ref["niter"] == single.convInfo["finIter"]
    assert np.allclose(list(multi.coef().values()), ref["par"], rtol=1e-8)
    assert multi.deviance() == pytest.approx(ref["ssr"], rel=1e-8)
    assert multi.nobs() == n
    assert np.allclose(multi.vcov(), single.vcov(), rtol=1e-8)
    assert np.allclose(multi.residuals(), single.residuals(), rtol=1e-9, atol=1e-12)
    assert "grad" not in multi.cfit and multi._resid is not None   # lazy: only residuals() has been asked for
    assert np.allclose(multi.gradient(), single.gradient(), rtol=1e-9, atol=1e-12)
    assert abs(np.sum(multi.residuals() ** 2) - multi.deviance()) <= 1e-9 * multi.deviance()
    # twice in a row: the device group and the per-device buffers are reused
    again = gsl_nls_large("y ~ A * exp(-lam * x) + b", devices=list(range(ngpu)), **kw)
    assert list(again.coef().values()) == list(multi.coef().values())


@pytest.mark.parametrize("ngpu", [1, 2, 4, 8])
def test_session_lazy_residuals_and_gradient(ngpu):
    """gslnls_session_*: the handle a host object keeps.  The fit returns no O(n) arrays; residuals and the
    n x p gradient (src/nls_large.c:339-385) come from the resident shards when asked, stitched to n rows."""
    if _ngpu() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    import bench
    import gslnls_b200 as G
    from oracle import oracle as O
    n = 300_007
    x, y = bench.synth_rows(0, n, n)
    w = 0.5 + (np.arange(n) % 5) / 4.0
    m = G.Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    for weights in (None, w):
        ses = G.Session(m, n, weights is not None, list(range(ngpu))).upload([x], y, weights)
        assert ses.ngpu == ngpu
        for alg in ("lm", "lmaccel"):
            fit = ses.fit(list(bench.START), algorithm=alg)
            ref = O.nls_large("exp3", y, list(bench.START), x=x, algorithm=alg, weights=weights, want_resid_grad=True)
            assert fit["conv"] == ref["conv"] == 0 and fit["niter"] == ref["niter"], alg
            assert np.allclose(fit["par"], ref["par"], rtol=1e-8) and fit["n"] == n
            assert "resid" not in fit
        r, J = ses.residuals(fit["par"], want_grad=True)
        assert np.allclose(r, ref["resid"], rtol=1e-6, atol=1e-8)
        assert np.allclose(J, ref["grad"], rtol=1e-6, atol=1e-8)
        assert abs(r @ r - fit["ssr"]) <= 1e-9 * fit["ssr"]
        ses.close()
    # the high-level call keeps the session behind the fitted object
    obj = G.gsl_nls_large("y ~ A * exp(-lam * x) + b", data={"x": x, "y": y}, start={"A": 1, "lam": 1, "b": 0},
                          jac=True, devices=list(range(max(ngpu, 2))) if ngpu > 1 else None)
    assert obj._resid is None                       # nothing O(n) was materialised by the fit
    assert abs(np.sum(obj.residuals() ** 2) - obj.deviance()) <= 1e-9 * obj.deviance()
    assert obj.gradient().shape == (n, 3)
