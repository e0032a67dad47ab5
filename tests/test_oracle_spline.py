"""The restated FITPACK evaluation (oracle/spline_oracle.py) against SciPy's own bispev on the same knots and
coefficients, for the splines the reference builds: cubic on a regular (x, t) grid (Burger / KS / Advection), linear
(Diffusion), queried at shifted / wrapped grid points including both domain ends."""
import numpy as np
import pytest

from oracle import spline_oracle as so


@pytest.mark.parametrize("kind", [3, 1])
def test_restated_bispev_matches_scipy(kind):
    from scipy.interpolate import RectBivariateSpline
    rng = np.random.default_rng(0)
    L, nx, nt = 2 * np.pi, 48, 21
    x = np.linspace(0, L, nx, endpoint=False)
    t = np.arange(nt) * 1e-2
    uu = np.sin(3 * x[None, :] + 5 * t[:, None]) + 0.1 * rng.normal(size=(nt, nx))
    spl = RectBivariateSpline(x, t, uu.T, kx=kind, ky=kind, s=0)
    tx, ty, c = spl.tck
    xq = np.concatenate((rng.uniform(0, L, 12), [0.0, x[-1], x[1], L, L + 0.3, -0.2]))      # incl. ends and out-of-domain
    tq = np.concatenate((rng.uniform(0, t[-1], 5), [0.0, t[-1], t[3]]))
    got = so.table(tx, ty, c, kind, kind, xq[None], tq)[0]
    xs = np.clip(xq, x[0], x[-1])
    want = np.array([[spl.ev(xv, tv) for xv in xs] for tv in tq])
    assert np.max(np.abs(got - want)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
    # on the data points an interpolating spline returns the data
    on = so.table(tx, ty, c, kind, kind, x[None, ::7], t[::5])[0]
    assert np.max(np.abs(on - uu[::5, ::7])) < 1e-12


@pytest.mark.parametrize("k,mx,mt", [(3, 32, 21), (1, 16, 9), (3, 9, 8), (3, 64, 101)])
def test_restated_fit_matches_scipy(k, mx, mt):
    """oracle.spline_oracle.fit (FITPACK regrid with s = 0: knot placement of fpregr + the collocation solve) against
    SciPy's RectBivariateSpline -- the stand-in for interp2d(x, t, uu, kind) of setGroundTruth (Burger.py:322-323)."""
    from scipy.interpolate import RectBivariateSpline
    from oracle import spline_oracle as so
    rng = np.random.default_rng(0)
    x = np.linspace(0, 2 * np.pi, mx, endpoint=False)
    t = np.arange(mt) * 1e-3
    z = np.sin(x[None, :] + 3 * t[:, None]) + 0.1 * rng.normal(size=(mt, mx))
    tx, ty, c = so.fit(x, t, z, k)
    TX, TY, C = RectBivariateSpline(x, t, z.T, kx=k, ky=k, s=0).tck
    assert np.array_equal(tx, TX) and np.array_equal(ty, TY)
    assert np.max(np.abs(c - C)) <= 1e-13 * np.max(np.abs(C))
