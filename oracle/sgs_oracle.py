"""TEST INFRASTRUCTURE ONLY.  Numpy restatement of the a-priori sub-grid-scale diagnostic of the reference,
``Burger.compute_Sgs(nURG)`` (/root/reference/python/_model/Burger.py:677-736) and ``KS.compute_Sgs`` (KS.py:385-409),
vectorised over the history rows.  Pinned by tests/golden/sgs.npz (recorded from the real classes,
tests/golden/make_golden_sgs.py)."""
import numpy as np
from scipy.fftpack import fft, ifft


def compute_sgs(uu, k, dx, dt, nu, nURG, ks=False):
    """uu [rows, N] real history, k [N] dimensional wavenumbers (FFT order).  Returns sgs [, alt, alt2]."""
    rows, N = uu.shape
    cut = np.abs(k) > nURG // 2                                  # :678
    r = nURG / N

    def filt(a):
        v = fft(a, axis=-1)
        v[:, cut] = 0                                            # vh aliases v: the spectrum itself is filtered (:691-693)
        return v, np.real(ifft(v, axis=-1))

    v, uh = filt(uu)
    _, u2h = filt(uu * uu)
    duhdx = (uh - np.roll(uh, 1, axis=-1)) / dx                  # :727
    du2hdx = (u2h - np.roll(u2h, 1, axis=-1)) / dx               # :730
    sgs = -uh * duhdx + 0.5 * du2hdx                             # :734
    if ks:
        return sgs
    nxt = np.r_[np.arange(1, rows), rows - 2]                    # :686: the last row looks back, with the sign flipped (:712-714)
    sign = np.ones((rows, 1))
    sign[-1] = -1
    vpt, uhpt = v[nxt], uh[nxt]
    d2 = (np.roll(uh, -1, axis=-1) - 2.0 * uh + np.roll(uh, 1, axis=-1)) / dx ** 2
    alt = sign * ((uhpt - uh) / dt) + uh * duhdx - nu * d2       # :735

    def coarse(w):                                               # :695, :708
        return np.real(ifft(np.concatenate((w[:, :(nURG + 1) // 2], w[:, -(nURG - 1) // 2:]), axis=-1), axis=-1)) * r

    a0, a1 = coarse(v), coarse(vpt)
    dudt2 = sign * ((a1 - a0) / dt)
    dudx2 = (a0 - np.roll(a0, 1, axis=-1)) / dx * r              # :732
    d22 = (np.roll(a0, -1, axis=-1) - 2.0 * a0 + np.roll(a0, 1, axis=-1)) / dx ** 2 * r ** 2
    alt2 = dudt2 + a0 * dudx2 - nu * d22                          # :736
    return sgs, alt, alt2
