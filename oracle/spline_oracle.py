"""TEST INFRASTRUCTURE ONLY (imported by tests/ only; the product path never touches oracle/).

CPU restatement of the tensor-product B-spline evaluation behind the reference's ground truth
(``scipy.interpolate.interp2d(x, t, uu, kind='cubic')``, /root/reference/python/_model/Burger.py:322-327, 578-589;
KS.py:221-223; Diffusion.py:130 with kind='linear').  The algorithm lives in a third-party dependency that is not
under /root/reference: FITPACK (P. Dierckx) as shipped by SciPy (reference pin scipy==1.6.2,
python/requirements.txt:5; SciPy here: 1.18): routine ``bispev`` -> ``fpbisp`` (knot-interval search with the
argument clamped to the spline's domain) and ``fpbspl`` (de Boor's recurrence for the k+1 non-zero B-splines).
Pinned by tests/test_oracle_spline.py against SciPy's own evaluation of the same (knots, coefficients).
The CUDA kernel csrc/spline_launch.cu follows this file line by line.
"""
import numpy as np


def find_span(t, k, x):
    """Index l with t[l] <= x < t[l+1], clamped to k .. len(t)-k-2 (fpbisp: the last interval takes x == t_end)."""
    lo, hi = k, len(t) - k - 1
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        if x >= t[mid]:
            lo = mid
        else:
            hi = mid
    return lo


def bspl(t, k, x, l):
    """fpbspl: the k+1 non-zero B-splines of degree k at x in the knot interval l."""
    h = np.zeros(k + 1)
    h[0] = 1.0
    for j in range(1, k + 1):
        hh = h[:j].copy()
        h[0] = 0.0
        for i in range(1, j + 1):
            li, lj = l + i, l + i - j
            if t[li] == t[lj]:
                h[i] = 0.0
                continue
            f = hh[i - 1] / (t[li] - t[lj])
            h[i - 1] = h[i - 1] + f * (t[li] - x)
            h[i] = f * (x - t[lj])
    return h


def bispev(tx, ty, c, kx, ky, x, y):
    """S(x, y) for scalar x, y."""
    x = min(max(x, tx[kx]), tx[len(tx) - kx - 1])
    y = min(max(y, ty[ky]), ty[len(ty) - ky - 1])
    lx, ly = find_span(tx, kx, x), find_span(ty, ky, y)
    hx, hy = bspl(tx, kx, x, lx), bspl(ty, ky, y, ly)
    ncy = len(ty) - ky - 1
    sp = 0.0
    for a in range(kx + 1):
        base = (lx - kx + a) * ncy + (ly - ky)
        for b in range(ky + 1):
            sp += c[base + b] * hx[a] * hy[b]
    return sp


def table(tx, ty, c, kx, ky, xq, tq):
    """out[q, i, j] = S(xq[q, j], tq[i])."""
    xq = np.atleast_2d(xq)
    out = np.empty((xq.shape[0], len(tq), xq.shape[1]))
    for q in range(xq.shape[0]):
        for i, tv in enumerate(tq):
            for j, xv in enumerate(xq[q]):
                out[q, i, j] = bispev(tx, ty, c, kx, ky, xv, tv)
    return out


def fit_knots(g, k):
    """FITPACK fpregr, s = 0, odd degree k: [g_0 (k+1 times), g_{k/2+1} .. g_{m-2-k/2}, g_{m-1} (k+1 times)]."""
    g = np.asarray(g, dtype=np.float64)
    m, k3 = len(g), k // 2
    return np.concatenate((np.full(k + 1, g[0]), g[k3 + 1:m - 1 - k3], np.full(k + 1, g[-1])))


def collocation_band(g, t, k):
    """ab[i, j - i + 3] = B_j(g_i): the interpolation matrix of the knot vector t in band storage (|j - i| <= 3)."""
    m = len(g)
    ab = np.zeros((m, 7))
    for i, xv in enumerate(g):
        l = find_span(t, k, xv)
        h = bspl(t, k, xv, l)
        for a in range(k + 1):
            j = l - k + a
            if 0 <= j - i + 3 < 7:
                ab[i, j - i + 3] = h[a]
    return ab


def band_lu(ab):
    """In-place LU without pivoting of a band matrix with half-widths 3 (the collocation matrix is totally positive)."""
    m = len(ab)
    for i in range(m):
        piv = ab[i, 3]
        for r in range(i + 1, min(i + 4, m)):
            mlt = ab[r, i - r + 3] / piv
            ab[r, i - r + 3] = mlt
            for j in range(i + 1, min(i + 4, m)):
                ab[r, j - r + 3] -= mlt * ab[i, j - i + 3]
    return ab


def band_solve(ab, rhs):
    """rhs [m, nrhs] -> solution, with the factors of band_lu."""
    m = len(ab)
    y = np.array(rhs, dtype=np.float64)
    for i in range(1, m):
        for j in range(max(i - 3, 0), i):
            y[i] -= ab[i, j - i + 3] * y[j]
    for i in range(m - 1, -1, -1):
        for j in range(i + 1, min(i + 4, m)):
            y[i] -= ab[i, j - i + 3] * y[j]
        y[i] /= ab[i, 3]
    return y


def fit(x, t, z, k):
    """Interpolating tensor-product spline through z[it, ix] (what interp2d / RectBivariateSpline(s=0) build; FITPACK regrid):
    returns (tx, ty, c) with c[ix * len(t) + it].  csrc/spline_launch.cu (mpde_fit_spline) follows this function."""
    tx, ty = fit_knots(x, k), fit_knots(t, k)
    abx, aby = band_lu(collocation_band(x, tx, k)), band_lu(collocation_band(t, ty, k))
    w = band_solve(abx, np.asarray(z, dtype=np.float64).T)            # [mx, mt]: along x for every time row
    c = band_solve(aby, w.T).T                                         # along t for every x coefficient -> [mx, mt]
    return tx, ty, np.ascontiguousarray(c).reshape(-1)
