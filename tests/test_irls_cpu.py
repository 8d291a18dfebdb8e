"""IRLS robust losses without a GPU: the oracle's restatement (oracle/irls.py) of src/nls_irls.c, checked by the
defining properties of the eight psi families and on a contaminated Example-1 data set."""
import numpy as np
import pytest

from oracle import irls as OI
from oracle import oracle as O


@pytest.mark.parametrize("loss", sorted(OI.LOSSES))
def test_psi_families_have_their_defining_properties(loss):
    which, c = OI.LOSSES[loss], OI.CC_DEFAULT[loss] + [0.0, 0.0]
    xs = np.concatenate([np.linspace(-12, 12, 481), [1e-3, -1e-3]])
    for x in xs:
        ps, dps = OI.psi(float(x), c, which)
        pm, _ = OI.psi(float(-x), c, which)
        assert ps == pytest.approx(-pm, abs=1e-15)                       # odd
        h = 1e-6
        num = (OI.psi(x + h, c, which)[0] - OI.psi(x - h, c, which)[0]) / (2 * h)
        kink = min(abs(abs(x) - k) for k in ([c[0], 1.5 * c[0], 3.5 * c[0], 8 * c[0], c[2], c[1], c[0] + c[1], 2 * c[0], 3 * c[0]])) < 1e-3
        if not kink and loss != "lqq":
            assert dps == pytest.approx(num, rel=1e-4, abs=1e-5), (loss, x)  # psi' really is d psi / dx
    slope0 = OI.psi(1e-3, c, which)[0] / 1e-3
    assert slope0 == pytest.approx(1.0 / (c[1] ** 2) if loss == "barron" else 1.0, rel=1e-4)  # psi(x) ~ x at 0
    if loss in ("bisquare", "optimal", "hampel"):
        assert OI.psi(50.0, c, which)[0] == 0.0                              # redescending to exactly 0
    if loss == "huber":
        assert OI.psi(50.0, c, which)[0] == c[0]


def test_median_matches_reference_definition():
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 6, 101, 1000):
        a = rng.standard_normal(n)
        assert OI.median(a) == np.median(a)


def test_irls_downweights_outliers(readme_examples):
    e = readme_examples["example1"]
    x, y = np.array(e["x"]), np.array(e["y"]).copy()
    y[[3, 11, 19]] += [6.0, -5.0, 8.0]                                      # gross outliers
    rows = lambda th: th[0] * np.exp(-th[1] * x) + th[2]  # noqa: E731
    ls = O.nls_large("exp3", y, [1.0, 1.0, 0.0], x=x)
    truth = np.array([5.0, 1.5, 1.0])
    for loss in ("huber", "bisquare", "welsh", "hampel", "lqq"):
        fit, info = OI.irls("exp3", y, [1.0, 1.0, 0.0], loss=loss, x=x, rows=rows)
        assert info["status"] == 0 and 2 <= info["niter"] <= 50, (loss, info)
        w = info["weights"]
        assert w[[3, 11, 19]].max() < 0.5 * np.median(w), loss               # outliers down-weighted
        assert abs(np.sum(w) - x.size) < 1e-9 * x.size                       # normalised to sum n
        assert np.linalg.norm(fit["par"] - truth) < np.linalg.norm(ls["par"] - truth), loss
