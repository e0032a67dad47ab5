"""Batched numpy restatement of the reference finite-difference environments.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
/root/reference/python/_model/Diffusion.py (IC :98-128, FDstep :137-162, step :164-216,
rewards/state :238-303) and Advection.py (IC :97-129, FDstep :138-152, step :154-213,
getMseReward :235-270, getState :272-286, analytic :289-291).

The reference builds a dense N x N matrix M and evaluates ``M @ u``; every row of M
has at most three non-zeros (sub-diagonal, diagonal, super-diagonal, periodic), so the
oracle forms the same three products per row and adds them in column order (the
remaining N-3 products are exact zeros).  BLAS may associate the three terms
differently, so agreement with the reference is to an ulp, not bitwise.
"""
import numpy as np

from .common import grid


def _tri_matvec(lo, di, up, u):
    """(M u)_k = lo_k u_{k-1} + di_k u_k + up_k u_{k+1} with periodic wrap, terms added in
    ascending column order as a dense row-times-vector product does."""
    um, upv = np.roll(u, 1, axis=-1), np.roll(u, -1, axis=-1)
    a, b, c = lo * um, di * u, up * upv
    out = (a + b) + c                       # interior rows: columns k-1, k, k+1
    out[..., 0] = (b[..., 0] + c[..., 0]) + a[..., 0]        # row 0: columns 0, 1, N-1
    out[..., -1] = (c[..., -1] + a[..., -1]) + b[..., -1]    # row N-1: columns 0, N-2, N-1
    return out


def sinus_ic(x, L, offset):
    """Diffusion.py:108 / Advection.py:108."""
    return np.sin((x - offset) * 2 * np.pi / L)


def box_ic(x, L, offset):
    """Diffusion.py:103-104."""
    u0 = np.zeros_like(x)
    u0[np.abs(x - L / 2 - offset) < L / 8] = 1.0
    return u0


def gaussian_ic(x, L, offset):
    """Diffusion.py:112."""
    return np.exp(-0.5 * (0.5 * L + offset - x) ** 2)


class DiffusionOracle:
    def __init__(self, B=1, L=2 * np.pi, N=32, dt=1e-3, nu=0.01, implicit=False):
        self.B, self.L, self.N, self.dt, self.nu = B, float(L), N, float(dt), nu
        self.dx = L / N
        self.x = grid(L, N)
        self.implicit = implicit

    def IC(self, u0):
        self.u0 = np.array(np.broadcast_to(u0, (self.B, self.N)), dtype=np.float64)
        self.u = self.u0.copy()
        self.t = 0.0
        self.ioutnum = 0

    def step(self, actions=None):
        """Diffusion.py:164-216.  ``actions``: None, [B, 1] (one global stencil weight) or
        [B, N] (per-gridpoint weights, MARL lists flattened)."""
        u = self.u
        if actions is None:
            if self.implicit:                                # :142-149
                c = self.dt * self.nu / self.dx ** 2
                M = np.diag(np.full(self.N, 1 + 2 * c)) + np.diag(np.full(self.N - 1, -c), 1) \
                    + np.diag(np.full(self.N - 1, -c), -1)
                M[0, -1] = -c
                M[-1, 0] = -c
                self.u = np.linalg.solve(M, u.T).T
            else:                                            # :156-160
                d2 = (-2.0 * u + np.roll(u, 1, axis=-1) + np.roll(u, -1, axis=-1)) / self.dx ** 2
                self.u = u + self.dt * self.nu * d2
        else:
            a = np.asarray(actions, dtype=np.float64).reshape(self.B, -1)
            if a.shape[1] == 1:
                a = np.broadcast_to(a, (self.B, self.N))     # :172-178 same stencil everywhere
            d2 = _tri_matvec(-a / 2, a, -a / 2, u)           # :190-202
            self.grad_last = d2
            self.u = u + self.dt * self.nu * d2 / self.dx ** 2   # :206
        self.t += self.dt
        self.ioutnum += 1

    def analytic(self, t=None):
        """Diffusion.py:301-303 (sinus case)."""
        t = self.t if t is None else t
        return self.u0 * np.exp(-(2.0 * np.pi / self.L) ** 2 * self.nu * t)

    def mse_reward(self, truth_row, numAgents=1):
        """Diffusion.py:245-252: -mean((truth - u)^2) per agent section -> [B, A]."""
        d = (truth_row - self.u) ** 2
        return -d.reshape(self.B, numAgents, self.N // numAgents).mean(axis=-1)

    def direct_reward(self):
        """Diffusion.py:275-281 -> [B, N]."""
        u = self.u
        d2 = (-2.0 * u + np.roll(u, 1, axis=-1) + np.roll(u, -1, axis=-1)) / self.dx ** 2
        return -np.power(d2, 2) / self.N

    def state(self, numAgents=1):
        """Diffusion.py:284-298: u, or windows uext[i*sec : (i+1)*sec + 2] -> [B, A, sec+2]."""
        if numAgents == 1:
            return self.u
        sec = self.N // numAgents
        ext = np.concatenate((self.u[:, -1:], self.u, self.u[:, :1]), axis=-1)
        return np.stack([ext[:, i * sec:(i + 1) * sec + 2] for i in range(numAgents)], axis=1)


class AdvectionOracle:
    def __init__(self, B=1, L=2 * np.pi, N=32, dt=1e-3, nu=0.01, offset=0.0):
        self.B, self.L, self.N, self.dt, self.nu = B, float(L), N, float(dt), nu
        self.dx = L / N
        self.x = grid(L, N)
        self.alpha = nu * dt / self.dx                       # Advection.py:43
        self.offset = offset

    def IC(self, u0):
        self.u0 = np.array(np.broadcast_to(u0, (self.B, self.N)), dtype=np.float64)
        self.u = self.u0.copy()
        self.t = 0.0
        self.ioutnum = 0

    def step(self, actions=None):
        """Advection.py:154-213.  ``actions``: None (Lax), [B, 2] (one global stencil:
        a0 multiplies u_{k-1}, a1 multiplies u_{k+1}) or [B, 2N] (per point: entry 2j
        multiplies u_{k+1}, entry 2j+1 multiplies u_{k-1} -- the opposite convention --
        except in the last row k = N-1, where the reference swaps them again)."""
        u = self.u
        one = np.ones((self.B, self.N))
        if actions is None:
            lo, di, up = (0.5 + 0.5 * self.alpha) * one, 0.0 * one, (0.5 - 0.5 * self.alpha) * one
        else:
            a = np.asarray(actions, dtype=np.float64).reshape(self.B, -1)
            if a.shape[1] == 2:
                lo, up = a[:, 0:1] * one, a[:, 1:2] * one
                di = (1 - (a[:, 0:1] + a[:, 1:2])) * one      # 1 - sum(actions) (:165)
            else:
                up, lo = a[:, 0::2].copy(), a[:, 1::2].copy()
                di = 1.0 - up - lo                            # (:182)
                # the last row swaps the convention: even entry -> u_{N-2}, odd -> u_0 (:188-190)
                up[:, -1], lo[:, -1] = a[:, -1], a[:, -2]
        self.u = _tri_matvec(lo, di, up, u)
        self.t += self.dt
        self.ioutnum += 1

    def analytic(self, t=None):
        """Advection.py:289-291."""
        t = self.t if t is None else t
        return np.sin((self.x - self.nu * t - self.offset) * 2 * np.pi / self.L)

    def mse_reward(self, numAgents=1):
        d = (self.analytic() - self.u) ** 2
        return -d.reshape(self.B, numAgents, self.N // numAgents).mean(axis=-1)

    def state(self, numAgents=1):
        if numAgents == 1:
            return self.u
        sec = self.N // numAgents
        ext = np.concatenate((self.u[:, -1:], self.u, self.u[:, :1]), axis=-1)
        return np.stack([ext[:, i * sec:(i + 1) * sec + 2] for i in range(numAgents)], axis=1)
