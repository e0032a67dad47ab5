"""TEST INFRASTRUCTURE ONLY.  Numpy restatements of the two remaining finite-difference environments of the reference
(no CUDA path yet -- these oracles and their goldens are the first step of that row):

* DiffusionError (/root/reference/python/_model/DiffusionError.py:150-200): the action is the ERROR of the stencil,
  row k of M = (1 - a_k/2, -2 + a_k, 1 - a_k/2); with one agent the two wrap-around entries are
  M[0,-1] = 1 - ac[0] = a/2 and M[-1,0] = 1 + ac[2] = 2 - a/2 (as written in the reference);
  u <- u + dt nu (M u) / dx^2.
* Laplace (/root/reference/python/_model/Laplace.py:105-148): N+1 points, agent i owns row i+1 with three free
  stencil entries, u <- u + dt (M u), Dirichlet point u[0] = 1; reward -(u_xx - force)^2 on points 1..N,
  state [u_{i-1}, u_i, u_{i+1}, force_i] for i < numAgents.

Dense matrix products are restated as the (at most) three products per row, added in ascending column order.
Pinned by tests/golden/fd_extra.npz (recorded from the real classes, tests/golden/make_golden_fd2.py)."""
import numpy as np

from .common import grid
from .fd_oracle import _tri_matvec


class DiffusionErrorOracle:
    def __init__(self, L=2 * np.pi, N=32, dt=1e-3, nu=0.01):
        self.L, self.N, self.dt, self.nu = float(L), N, float(dt), nu
        self.dx = L / N
        self.x = grid(L, N)

    def IC(self, u0):
        self.u0 = np.asarray(u0, dtype=np.float64).copy()
        self.u = self.u0.copy()
        self.t, self.ioutnum = 0.0, 0

    def step(self, actions=None, numAgents=1):
        N, u = self.N, self.u
        if actions is None:                                    # FDstep, explicit (:137-147)
            d2 = (-2.0 * u + np.roll(u, 1) + np.roll(u, -1)) / self.dx ** 2
            self.u = u + self.dt * self.nu * d2
        else:
            a = np.asarray(actions, dtype=np.float64).reshape(-1)
            if numAgents == 1:                                 # :153-160
                lo = np.full(N, 1 - a[0] / 2); di = np.full(N, -2 + a[0]); up = np.full(N, 1 - a[0] / 2)
                lo[0] = 1 - (1 - a[0] / 2)                     # M[0,-1]  = 1 - ac[0]
                up[-1] = 1 + (1 - a[0] / 2)                    # M[-1,0] = 1 + ac[2]
            else:                                              # :162-176
                lo = 1 - a / 2; di = -2 + a; up = 1 - a / 2
            d2 = _tri_matvec(lo[None], di[None], up[None], u[None])[0]
            self.u = u + self.dt * self.nu * d2 / self.dx ** 2  # :181
        self.t += self.dt
        self.ioutnum += 1


class LaplaceOracle:
    def __init__(self, L=2 * np.pi, N=32, dt=0.01):
        self.N = int(N) + 1                                    # Laplace.py:13
        self.L, self.dt = float(L), float(dt)
        self.dx = L / self.N
        self.x = np.linspace(0, self.L, self.N, endpoint=False)

    def IC(self, u0, force):
        self.u = np.asarray(u0, dtype=np.float64).copy()
        self.force = np.asarray(force, dtype=np.float64).copy()
        self.t, self.ioutnum = 0.0, 0

    def step(self, actions, numAgents):
        N, u = self.N, self.u
        assert numAgents + 1 == N
        a = np.asarray(actions, dtype=np.float64).reshape(numAgents, 3)
        d2 = np.zeros(N)
        for i in range(numAgents):                             # row i+1: columns i % N, i+1, (i+2) % N  (:109-113)
            cols = [(i % N, a[i, 0]), (i + 1, a[i, 1]), ((i + 2) % N, a[i, 2])]
            acc = 0.0
            for c, w in sorted(cols):                          # dense row times vector: ascending column order
                acc = acc + w * u[c]
            d2[i + 1] = acc
        self.u = u + self.dt * d2                              # :116
        self.u[0] = 1.0                                        # :118
        self.t += self.dt
        self.ioutnum += 1

    def direct_reward(self):
        u = self.u
        d2 = (-2.0 * u + np.roll(u, 1) + np.roll(u, -1)) / self.dx ** 2
        return -np.power(d2[1:] - self.force[1:], 2)           # :139-144

    def state(self, numAgents):
        N, u = self.N, self.u
        return np.array([[u[(i - 1) % N], u[i], u[(i + 1) % N], self.force[i]] for i in range(numAgents)])   # :146-150
