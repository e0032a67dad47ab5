"""Batched numpy restatement of the reference Burgers environment stepper.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
/root/reference/python/_model/Burger.py -- IC :205-320, step :333-499,
compute_Ek :541-576, getMseReward :578-601, getState :604-675 -- and the reward
bookkeeping of burger_environment.py:99-176.  Time integration is AB2 for the
nonlinear term + Crank-Nicolson for viscosity ("ABCN", Burger.py:486-491).

Fields are [B, N]; row e is exactly what the reference computes for its single
environment (numpy broadcasting only, same op order per row).
"""
import numpy as np

from .common import (fft, ifft, grid, wavenumbers, action_basis, laplacian_fd, upwind_fd,
                     energy_row_f32, RunningSpectrum, agent_windows)


def forcing_tables(seed, nsteps):
    """randfac1, randfac2 = normal((32, nsteps)) drawn right after np.random.seed(seed)
    (Burger.py:66, 94-95; no draw in between unless nunoise)."""
    rs = np.random.RandomState(seed)
    r1 = rs.normal(loc=0.0, scale=1.0, size=(32, nsteps))
    r2 = rs.normal(loc=0.0, scale=1.0, size=(32, nsteps))
    return r1, r2


def turbulence_ic(x, L, N, offset, tseed):
    """'turbulence' initial field (Burger.py:227-260): k^-5/3 spectrum, phases from the
    LCG rng <- (1103515245 rng + 12345) mod 2^13, rescaled until rms(u-1) in [0.65, 0.75]."""
    rng = 123456789 + int(tseed)
    u0 = np.ones(N)
    for k in range(1, N):
        rng = (1103515245 * rng + 12345) % 2 ** 13
        phase = rng / 2 ** 13 * 2.0 * np.pi
        Ek = 5 ** (-5 / 3) if k <= 5 else k ** (-5 / 3)
        u0 += np.sqrt(2 * Ek) * np.sin(k * 2 * np.pi * (x + offset) / L + phase)
    crit = np.sqrt(np.sum((u0 - 1.0) ** 2) / N)
    it = 0
    while crit < 0.65 or crit > 0.75:
        u0 *= 0.7 / crit
        crit = np.sqrt(np.sum((u0 - 1.0) ** 2) / N)
        it += 1
        if it > 100:
            break
    return u0


def sinus_ic(x, L, offset):
    """Burger.py:224."""
    return np.sin(4.0 * np.pi * (x + offset) / L)


class BurgerOracle:
    """State container + step for a batch of independent Burgers environments."""

    def __init__(self, B=1, L=2 * np.pi, N=32, dt=1e-3, nu=0.02, dforce=True, forcing=False,
                 ssm=False, dsm=False, stepper=1, version=0, numAgents=1, offset=0.0):
        assert not (ssm and dsm)                              # Burger.py:50
        self.B, self.L, self.N, self.dt = B, float(L), N, float(dt)
        self.nu = np.broadcast_to(np.asarray(nu, dtype=np.float64), (B,)).reshape(B, 1)
        self.dx = L / N                                       # Burger.py:85 (L as given)
        self.x = grid(L, N)
        self.k = wavenumbers(L, N)                            # Burger.py:161-163
        self.k1 = 1j * self.k
        self.k2 = self.k1 ** 2
        self.dforce, self.forcing, self.ssm, self.dsm = dforce, forcing, ssm, dsm
        self.stepper, self.version, self.numAgents = stepper, version, numAgents
        self.offset = np.broadcast_to(np.asarray(offset, dtype=np.float64), (B,)).reshape(B, 1)
        self.cs = 0.1                                         # Burger.py:104
        self.basis, self.M = None, 0
        self.randfac1 = self.randfac2 = None                  # [32, cols] shared or [B, 32, cols]
        self.truth = None                                     # callable (xq[B,N], t) -> [B,N]

    # -- set-up ---------------------------------------------------------------
    def setup_basis(self, M, kind="uniform"):
        self.M = M
        self.basis = action_basis(self.x, self.L, M, kind)

    def set_forcing_tables(self, r1, r2):
        self.randfac1, self.randfac2 = np.asarray(r1), np.asarray(r2)

    def IC(self, u0=None, v0=None):
        """Burger.py:289-320: u0 -> v0 = fft(u0); v0 -> u0 = Re ifft(v0) with v kept as given."""
        if v0 is None:
            u0 = np.array(np.broadcast_to(u0, (self.B, self.N)), dtype=np.float64)
            v0 = fft(u0, axis=-1)
        else:
            v0 = np.array(np.broadcast_to(v0, (self.B, self.N)), dtype=np.complex128)
            u0 = np.real(ifft(v0, axis=-1))
        self.u, self.v = u0, v0
        self.u_prev = u0.copy()                               # getState at ioutnum==0: umt=u (:609)
        self.t = 0.0
        self.ioutnum = 0
        self.Fn_old = self.k1 * fft(0.5 * self.u ** 2, axis=-1)  # Burger.py:320
        self.spec = RunningSpectrum(self.v, self.N, self.dx)  # row 0 of vv (:316)
        self.sgs_last = np.zeros_like(u0)

    # -- forcing pieces -------------------------------------------------------
    def _stochastic(self):
        """Burger.py:410-421."""
        f = np.zeros((self.B, self.N))
        A = np.sqrt(2.0) / self.L
        c = self.ioutnum % self.stepper
        for k in range(1, 4):
            r1 = self.randfac1[..., k, c].reshape(-1, 1)
            r2 = self.randfac2[..., k, c].reshape(-1, 1)
            f = f + r1 * A / np.sqrt(k * self.stepper * self.dt) * np.cos(
                2 * np.pi * k * (self.x + self.offset) / self.L + 2 * np.pi * r2)
        return f

    def _static_smagorinsky(self):
        """Burger.py:337-349.  delta = 2 pi / N regardless of L."""
        delta = 2 * np.pi / self.N
        return (self.cs * delta) ** 2 * np.abs(upwind_fd(self.u, self.dx)) * laplacian_fd(self.u, self.dx)

    def _dynamic_smagorinsky(self):
        """Burger.py:354-408.  The sharp test filter |k| > N//4 is applied IN PLACE to the
        live spectrum self.v (vh aliases self.v, :369-370) -- reproduced here."""
        N, dx = self.N, self.dx
        delta, deltah = 2 * np.pi / N, 4 * np.pi / N
        cut = np.abs(self.k) > N // 4
        v2 = fft(self.u ** 2, axis=-1)
        v2[:, cut] = 0
        L1 = 0.5 * np.real(ifft(v2, axis=-1))
        self.v = self.v.copy()
        self.v[:, cut] = 0
        uh = np.real(ifft(self.v, axis=-1))
        Lg = L1 - 0.5 * uh ** 2
        dudx = upwind_fd(self.u, dx)
        d2 = laplacian_fd(self.u, dx)
        w2 = fft(np.abs(dudx) * dudx, axis=-1)
        w2[:, cut] = 0
        M1 = delta ** 2 * np.real(ifft(w2, axis=-1))
        duh = upwind_fd(uh, dx)
        M2 = deltah ** 2 * np.abs(duh) * duh
        malt = 4.0 / deltah ** 2 * M2 - 1.0 / delta ** 2 * M1
        Malt = (malt - np.roll(malt, 1, axis=-1)) / dx
        c = np.mean(-Lg * Malt, axis=-1, keepdims=True) / np.mean(Malt * Malt, axis=-1, keepdims=True)
        return c * np.abs(dudx) * d2

    # -- one solver step --------------------------------------------------------
    def step(self, actions=None):
        """Burger.py:333-499.  ``actions`` is [B, M] (MARL lists already flattened, :437)."""
        B, N = self.B, self.N
        # Q1: the accumulator starts as complex64 and `+=` keeps that dtype (:335, 352, 408, 466)
        F = np.zeros((B, N), dtype=np.complex64)
        if self.ssm:
            self.sgs_last = self._static_smagorinsky()
            F += fft(self.sgs_last, axis=-1)
        if self.dsm:
            self.sgs_last = self._dynamic_smagorinsky()
            F += fft(self.sgs_last, axis=-1)
        if self.forcing:
            F = fft(self._stochastic(), axis=-1)              # replaces, complex128 (:421)
        if actions is not None:
            a = np.asarray(actions, dtype=np.float64).reshape(B, self.M)
            f = a @ self.basis                                # :442
            if not self.dforce:
                f = f * laplacian_fd(self.u, self.dx)         # :445-450 (uu[ioutnum] == u)
            self.sgs_last = f
            F += fft(f, axis=-1)                              # :466
        C = -0.5 * self.k2 * self.nu * self.dt                # :486
        Fn = self.k1 * fft(0.5 * self.u ** 2, axis=-1)        # :487
        self.v = ((1.0 - C) * self.v - 0.5 * self.dt * (3.0 * Fn - self.Fn_old) + self.dt * F) / (1.0 + C)
        self.Fn_old = Fn
        self.u_prev = self.u
        self.u = np.real(ifft(self.v, axis=-1))               # :491
        self.t += self.dt                                     # :494 (accumulated, Q9)
        self.ioutnum += 1
        self.spec.push(self.v)                                # vv[ioutnum] = complex64(v) (:498)

    # -- observables ------------------------------------------------------------
    def state(self):
        """getState (Burger.py:604-675) -> [B, S] (numAgents==1) or [B, A, W]."""
        u = self.u
        d2 = laplacian_fd(u, self.dx)
        dudt = (u - self.u_prev) / self.dt
        ver, A = self.version, self.numAgents
        ek = 0.5 * np.real(self.v.conj() * self.v / self.N) * self.dx   # :653
        if A == 1:
            if ver == 0:
                return d2
            if ver == 1:
                return np.concatenate((dudt, d2), axis=-1)
            if ver == 2:
                return np.concatenate((u, u ** 2), axis=-1)
            base = d2 if ver == 3 else u
            return np.concatenate((base, ek[:, :self.N // 2]), axis=-1)
        if ver == 0:
            return agent_windows(d2, A)
        if ver == 1:
            return np.concatenate((agent_windows(dudt, A), agent_windows(d2, A)), axis=-1)
        if ver == 2:
            return np.concatenate((agent_windows(u, A), agent_windows(u ** 2, A)), axis=-1)
        base = d2 if ver == 3 else u
        tail = np.broadcast_to(ek[:, None, :self.N // 2], (self.B, A, self.N // 2))
        return np.concatenate((agent_windows(base, A), tail), axis=-1)

    def Ek_ktt_row(self):
        """Row ``ioutnum`` of Ek_ktt (Burger.py:555) from the running float32 sum."""
        return self.spec.mean()

    def mse_reward(self, truth_row):
        """getMseReward (Burger.py:589-599) given the interpolated truth row [B, N]
        -> [B, A] = -mean over each agent's segment of (truth - u)^2."""
        d = (truth_row - self.u) ** 2
        A = self.numAgents
        return -d.reshape(self.B, A, self.N // A).mean(axis=-1)


def truncated_ic(dns_v0, dns_k, offset, gridSize):
    """LES IC by spectral truncation with phase shift (burger_environment.py:110-111).
    NOTE literal transcription: exp(1j * 2 pi * offset * k) with k the dimensional wavenumber."""
    from .common import truncate_spectrum
    return truncate_spectrum(dns_v0 * np.exp(1j * 2 * np.pi * offset * dns_k), gridSize)
