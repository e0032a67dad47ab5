"""Host-side (numpy) set-up math of the batched environments: grids, wavenumbers, action
bases, initial conditions, the spectrum of the stochastic forcing, ETDRK4 tables and the
ground-truth interpolant.  None of this is on the per-step path; it runs once per episode.
Reference lines are cited per function (paths relative to /root/reference/python/_model).
"""
import numpy as np


def grid_points(L, N):
    """Burger.py:86 -- N equispaced points on [0, L)."""
    return np.linspace(0, L, N, endpoint=False)


def fft_wavenumbers(L, N):
    """Burger.py:161 / KS.py:113 -- fftfreq(N, L/(2 pi N)): (2 pi/L) * [0..N/2-1, -N/2..-1]."""
    return np.fft.fftfreq(N, L / (2 * np.pi * N))


def make_basis(x, L, M, kind):
    """Burger.py:177-203 / KS.py:139-164 -- [M, N] map from actions to grid forcing."""
    N = len(x)
    if M > 1:
        if kind == 'uniform':
            assert N % M == 0, "[Burger] Something went wrong in basis setup"
            basis = np.kron(np.eye(M), np.ones(N // M))
        elif kind == 'hat':
            h = L / (M - 1)
            nodes = (np.arange(M) * h)[:, None]
            up = np.clip((x[None, :] + h - nodes) / h, a_min=0., a_max=1.)
            down = np.clip((h - x[None, :] + nodes) / h, a_min=0., a_max=1.)
            basis = up + down - 1.
        else:
            raise SystemExit("[Burger] Basis function not known, exit..")
    else:
        basis = np.ones((M, N))
    np.testing.assert_allclose(np.sum(basis, axis=0), 1)
    return basis


def turbulence_field(x, L, N, offset, tseed):
    """Burger.py:227-260 -- 1 + sum_k sqrt(2 E_k) sin(2 pi k (x+off)/L + phi_k), E_k = k^-5/3
    (flat below k = 5), phases from a 13-bit LCG seeded with 123456789 + tseed, rescaled to
    an rms fluctuation in [0.65, 0.75]."""
    state = 123456789 + int(tseed)
    phases = []
    for _ in range(N - 1):
        state = (1103515245 * state + 12345) % 8192
        phases.append(state / 8192 * 2. * np.pi)
    u0 = np.ones(N)
    for k, ph in enumerate(phases, start=1):      # python ints: k ** (-5/3) must go through libm pow
        Ek = 5 ** (-5 / 3) if k <= 5 else k ** (-5 / 3)
        u0 += np.sqrt(2 * Ek) * np.sin(k * 2 * np.pi * (x + offset) / L + ph)
    rms = np.sqrt(np.sum((u0 - 1.) ** 2) / N)
    tries = 0
    while rms < 0.65 or rms > 0.75:
        u0 *= 0.7 / rms
        rms = np.sqrt(np.sum((u0 - 1.) ** 2) / N)
        tries += 1
        if tries > 100:
            break
    assert 0.6 < rms < 0.8
    return u0


def forced_field(x, L, N, stream):
    """Burger.py:265-273 -- random sines drawn from the seeded stream AFTER the forcing tables."""
    u0 = np.zeros(N)
    A = 1. / N
    for k in range(1, N):
        r1 = stream.normal(loc=0., scale=1.)
        r2 = stream.normal(loc=0., scale=1.)
        u0 += r1 * A * np.sin(2. * np.pi * (k * x / L + r2))
    return u0


def forcing_spectrum_coefficients(r1, r2, offset, L, dt, stepper, N, B):
    """Spectrum of the stochastic forcing of Burger.py:410-421 at the only modes it excites.

    f_j = sum_{k=1..3} r1[k,c] * A / sqrt(k s dt) * cos(2 pi k (x_j + off)/L + 2 pi r2[k,c]),  A = sqrt(2)/L,
    so fft(f)[k] = (N/2) * amp_k * exp(i (2 pi k off / L + 2 pi r2[k,c])) for k = 1,2,3 (and the conjugate at
    -k), zero elsewhere (exact for N >= 8).  Returns float64 [n, stepper, 3, 2] (re, im), n = 1 when the
    tables and offsets are shared by all environments, else n = B.
    """
    r1, r2 = np.asarray(r1, dtype=np.float64), np.asarray(r2, dtype=np.float64)
    offset = np.asarray(offset, dtype=np.float64).reshape(-1)
    shared = r1.ndim == 2 and np.all(offset == offset[0])
    n = 1 if shared else B
    if r1.shape[-1] < stepper:
        raise ValueError("forcing tables need at least `stepper` columns")
    k = np.arange(1, 4, dtype=np.float64)
    a1 = r1[..., 1:4, :stepper]                       # [..., 3, s]
    a2 = r2[..., 1:4, :stepper]
    if a1.ndim == 2:
        a1, a2 = a1[None], a2[None]
    off = offset[:n, None, None]
    amp = a1 * (np.sqrt(2.) / L) / np.sqrt(k[None, :, None] * stepper * dt)
    phase = 2 * np.pi * k[None, :, None] * off / L + 2 * np.pi * a2
    coef = (N / 2.) * amp * np.exp(1j * phase)        # [n, 3, s]
    coef = np.broadcast_to(coef, (n, 3, stepper)).transpose(0, 2, 1)
    out = np.empty((n, stepper, 3, 2))
    out[..., 0], out[..., 1] = coef.real, coef.imag
    return np.ascontiguousarray(out)


def etdrk4_coefficients(L, N, dt):
    """KS.py:112-137 -- Kassam-Trefethen ETDRK4 tables with 62-point contour means.
    The reference ignores nu (linear operator k^2 - k^4)."""
    k = fft_wavenumbers(L, N)
    lin = k ** 2 - k ** 4
    E, E2 = np.exp(dt * lin), np.exp(dt * lin / 2.)
    MM = 62
    roots = np.exp(1j * np.pi * (np.r_[1:MM + 1] - 0.5) / MM)
    LR = dt * np.repeat(lin[:, np.newaxis], MM, axis=1) + np.repeat(roots[np.newaxis, :], N, axis=0)
    Q = dt * np.real(np.mean((np.exp(LR / 2.) - 1.) / LR, 1))
    f1 = dt * np.real(np.mean((-4. - LR + np.exp(LR) * (4. - 3. * LR + LR ** 2)) / (LR ** 3), 1))
    f2 = dt * np.real(np.mean((2. + LR + np.exp(LR) * (-2. + LR)) / (LR ** 3), 1))
    f3 = dt * np.real(np.mean((-4. - 3. * LR - LR ** 2 + np.exp(LR) * (4. - LR)) / (LR ** 3), 1))
    return dict(k=k, l=lin, E=E, E2=E2, Q=Q, f1=f1, f2=f2, f3=f3, g=-0.5j * k)


class TruthInterpolant:
    """Stand-in for ``scipy.interpolate.interp2d(x, t, uu, kind)`` (Burger.py:323, KS.py:223,
    Diffusion.py:132), which SciPy >= 1.14 no longer ships: the same FITPACK interpolating
    tensor-product spline through RectBivariateSpline, built lazily.  Calling it follows
    interp2d's semantics (inputs sorted, result [len(t), len(x)], 1-D for scalar t)."""

    def __init__(self, x, t, uu, kind='cubic'):
        self.x, self.t, self.uu = np.asarray(x, float), np.asarray(t, float), np.asarray(uu, float)
        self.order = {'linear': 1, 'cubic': 3}[kind]
        self._spl = None

    def _spline(self):
        if self._spl is None:
            from scipy.interpolate import RectBivariateSpline
            self._spl = RectBivariateSpline(self.x, self.t, self.uu.T, kx=self.order, ky=self.order, s=0)
        return self._spl

    def __call__(self, x, t):
        xs = np.sort(np.atleast_1d(np.asarray(x, float)))
        ts = np.sort(np.atleast_1d(np.asarray(t, float)))
        out = self._spline()(xs, ts).T
        return out[0] if out.shape[0] == 1 else out

    def fit_device(self, device):
        """Knots and coefficients of the interpolating spline computed ON THE DEVICE (mpde_fit_spline: FITPACK regrid with
        s = 0 restated -- same knots, collocation system solved by banded LU), cached per device.  Replaces the host fit
        (SciPy FITPACK, 0.26 - 0.48 s for a 5001 x 512 DNS) on the MSE-reward path."""
        import torch
        from . import _lib as LB
        if getattr(self, "_tck_dev", None) is not None and self._tck_dev[0] == str(device):
            return self._tck_dev
        lib = LB.lib()
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device)
        x, t, z = up(self.x), up(self.t), up(self.uu)
        mx, mt, k = x.numel(), t.numel(), self.order
        assert z.shape == (mt, mx), "truth array must be [len(t), len(x)]"
        tx = torch.empty(mx + k + 1, dtype=torch.float64, device=device)
        ty = torch.empty(mt + k + 1, dtype=torch.float64, device=device)
        c = torch.empty(mx * mt, dtype=torch.float64, device=device)
        work = torch.empty(mt * mx + 7 * (mx + mt), dtype=torch.float64, device=device)
        with torch.cuda.device(device):
            rc = lib.mpde_fit_spline(x.data_ptr(), mx, t.data_ptr(), mt, z.data_ptr(), k, tx.data_ptr(), ty.data_ptr(), c.data_ptr(),
                                     work.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        if rc != 0:
            raise RuntimeError("marlpde_b200: mpde_fit_spline failed (degree 1 or 3, at least 2 (k + 1) points per axis)")
        self._tck_dev = (str(device), tx, ty, c)
        return self._tck_dev

    def rows_device(self, xq, ts, device, dtype):
        """Truth tables on the GPU: xq [nq, N] (one shifted / wrapped grid per row), ts [rows] -> tensor [nq, rows, N].
        The spline is fitted on the device (``fit_device``) and sampled by ``mpde_eval_spline_table``."""
        import ctypes as C
        import torch
        from . import _lib as LB
        lib = LB.lib()
        _, tx, ty, c = self.fit_device(device)
        xq = np.atleast_2d(np.asarray(xq, dtype=np.float64))
        xd = torch.as_tensor(np.ascontiguousarray(xq), device=device)
        td = torch.as_tensor(np.ascontiguousarray(np.asarray(ts, dtype=np.float64)), device=device)
        out = torch.empty((xq.shape[0], td.numel(), xq.shape[1]), device=device, dtype=dtype)
        st = torch.cuda.current_stream(device).cuda_stream
        rc = lib.mpde_eval_spline_table(tx.data_ptr(), tx.numel(), ty.data_ptr(), ty.numel(), c.data_ptr(), self.order, self.order,
                                        xd.data_ptr(), xq.shape[0], xq.shape[1], td.data_ptr(), td.numel(), out.data_ptr(),
                                        LB.F64 if dtype == torch.float64 else LB.F32, st)
        if rc != 0:
            raise RuntimeError("marlpde_b200: mpde_eval_spline_table failed")
        return out

    def rows(self, xq, ts):
        """Truth at the (unsorted) points xq for every time in ts -> [len(ts), len(xq)]."""
        order = np.argsort(xq, kind='stable')
        vals = self._spline()(np.asarray(xq)[order], np.asarray(ts, float))   # [nx, nt]
        out = np.empty((len(ts), len(xq)))
        out[:, order] = vals.T
        return out
