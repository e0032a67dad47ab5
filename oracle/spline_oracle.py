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
