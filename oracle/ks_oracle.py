"""Batched numpy restatement of the reference Kuramoto-Sivashinsky stepper.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
/root/reference/python/_model/KS.py: Fourier/ETDRK4 tables :112-137, IC :166-219,
step :230-274, fou2real :316-320, compute_Ek :322-343, getState :369-383.
ETDRK4 after Kassam & Trefethen (SISC 2005), 62-point contour means.
"""
import numpy as np

from .common import fft, ifft, grid, wavenumbers, action_basis, laplacian_fd, RunningSpectrum


def etdrk4_tables(L, N, dt):
    """KS.py:117, 127-137.  NOTE nu is ignored by the reference (l = k^2 - k^4)."""
    k = wavenumbers(L, N)
    lin = k ** 2 - k ** 4
    E = np.exp(dt * lin)
    E2 = np.exp(dt * lin / 2.0)
    MM = 62
    r = np.exp(1j * np.pi * (np.arange(1, MM + 1) - 0.5) / MM)
    LR = dt * lin[:, None] + r[None, :]
    Q = dt * np.real(np.mean((np.exp(LR / 2.0) - 1.0) / LR, axis=1))
    f1 = dt * np.real(np.mean((-4.0 - LR + np.exp(LR) * (4.0 - 3.0 * LR + LR ** 2)) / LR ** 3, axis=1))
    f2 = dt * np.real(np.mean((2.0 + LR + np.exp(LR) * (-2.0 + LR)) / LR ** 3, axis=1))
    f3 = dt * np.real(np.mean((-4.0 - 3.0 * LR - LR ** 2 + np.exp(LR) * (4.0 - LR)) / LR ** 3, axis=1))
    g = -0.5j * k
    return dict(k=k, E=E, E2=E2, Q=Q, f1=f1, f2=f2, f3=f3, g=g)


class KSOracle:
    def __init__(self, B=1, L=22.0, N=64, dt=0.25, dforce=True):
        self.B, self.L, self.N, self.dt = B, float(L), N, dt
        self.dx = L / N
        self.x = grid(L, N)
        self.tab = etdrk4_tables(L, N, dt)
        self.dforce = dforce
        self.basis, self.M = None, 0

    def setup_basis(self, M, kind="uniform"):
        self.M = M
        self.basis = action_basis(self.x, self.L, M, kind)

    def IC(self, u0=None, v0=None):
        """KS.py:191-219."""
        if v0 is None:
            u0 = np.array(np.broadcast_to(u0, (self.B, self.N)), dtype=np.float64)
            v0 = fft(u0, axis=-1)
        else:
            v0 = np.array(np.broadcast_to(v0, (self.B, self.N)), dtype=np.complex128)
        self.v = v0
        self.t = 0.0
        self.ioutnum = 0
        self.spec = RunningSpectrum(self.v, self.N, self.dx)
        # dforce=False reads uu[ioutnum] (KS.py:241).  uu is only filled by fou2real
        # (KS.py:316-320): after a refresh at step s, row s is valid and rows > s are
        # zero; before the first refresh uu is complex64 and the reference raises a
        # casting error.  Modelled as: row valid iff refreshed at the current ioutnum.
        self.uu_row = None
        self.uu_valid_at = -1

    def _nonlinear(self, w):
        """N(w) = g * fft(Re(ifft(w))^2)  (KS.py:256-262)."""
        return self.tab["g"] * fft(np.real(ifft(w, axis=-1)) ** 2, axis=-1)

    def step(self, actions=None):
        """KS.py:230-274.  The forcing enters only the final combination."""
        T = self.tab
        F = None
        if actions is not None:
            a = np.asarray(actions, dtype=np.float64).reshape(self.B, self.M)
            f = a @ self.basis
            if not self.dforce:
                if self.uu_valid_at < 0:
                    raise TypeError("reference raises here: uu is still complex64 (call state() first)")
                row = self.uu_row if self.uu_valid_at == self.ioutnum else np.zeros_like(self.uu_row)
                um, up = np.roll(row, -1, axis=-1), np.roll(row, 1, axis=-1)   # float32 rolls (:242-243)
                f = f * ((up - 2.0 * row + um) / self.dx ** 2)                  # float32 stencil (:244)
            F = fft(f, axis=-1)
        v = self.v
        Nv = self._nonlinear(v)
        a_ = T["E2"] * v + T["Q"] * Nv
        Na = self._nonlinear(a_)
        b_ = T["E2"] * v + T["Q"] * Na
        Nb = self._nonlinear(b_)
        c_ = T["E2"] * a_ + T["Q"] * (2.0 * Nb - Nv)
        Nc = self._nonlinear(c_)
        if F is not None:
            self.v = T["E"] * v + (Nv + F) * T["f1"] + 2.0 * (Na + Nb + 2 * F) * T["f2"] + (Nc + F) * T["f3"]
        else:
            self.v = T["E"] * v + Nv * T["f1"] + 2.0 * (Na + Nb) * T["f2"] + Nc * T["f3"]
        self.t += self.dt
        self.ioutnum += 1
        self.spec.push(self.v)

    def fou2real_row(self):
        """Row ``ioutnum`` of uu after fou2real (KS.py:316-320): Re ifft of the COMPLEX64
        history row -> float32 (Q7)."""
        self.uu_row = np.real(ifft(self.v.astype(np.complex64), axis=-1))
        self.uu_valid_at = self.ioutnum
        return self.uu_row

    def state(self):
        """KS.py:369-383 on the float32 row: [dudx_central ; d2udx2] -> [B, 2N]."""
        u = self.fou2real_row()
        up, um = np.roll(u, -1, axis=-1), np.roll(u, 1, axis=-1)
        dudx = (up - um) / (2.0 * self.dx)
        d2 = (up - 2.0 * u + um) / self.dx ** 2
        return np.concatenate((dudx, d2), axis=-1)

    def Ek_ktt_row(self):
        return self.spec.mean()
