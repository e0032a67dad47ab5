"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's finite-difference Burgers solver
(/root/reference/python/_model/Burger_fd.py:335-476): explicit Euler in time, first-order upwind du/dx and centred
d2u/dx2 in space, the same closures / forcing / action handling as Burger.step, and v = fft(u) refreshed every step.
Pinned by tests/golden/burger_fd.npz (recorded from the real class, tests/golden/make_golden_fd.py)."""
import numpy as np
from scipy.fftpack import fft

from .burger_oracle import BurgerOracle
from .common import laplacian_fd, upwind_fd


class BurgerFdOracle(BurgerOracle):
    def __init__(self, *a, ssmforce=False, **kw):
        super().__init__(*a, **kw)
        self.ssmforce = ssmforce

    def step(self, actions=None):
        """Burger_fd.py:335-476."""
        B, N, dx = self.B, self.N, self.dx
        u = self.u
        forcing = np.zeros((B, N))
        if self.ssm:                                            # :343-355
            forcing = self._static_smagorinsky()
        if self.dsm:                                            # :358-411: the same closure as Burger.step (filters self.v in place;
            forcing = self._dynamic_smagorinsky()               # v is refreshed from u below, so nothing of it survives the step)
        if self.forcing:                                        # :406-417 (replaces the closure)
            forcing = self._stochastic()
        if actions is not None:                                 # :431-458
            a = np.asarray(actions, dtype=np.float64).reshape(B, self.M)
            af = a @ self.basis
            if not self.dforce:
                af = af * laplacian_fd(u, dx)
            if self.ssmforce:
                delta = 2 * np.pi / N
                af = (af * delta) ** 2 * np.abs(upwind_fd(u, dx)) * laplacian_fd(u, dx)
            forcing = forcing + af
        dudx, d2 = upwind_fd(u, dx), laplacian_fd(u, dx)        # :465-466
        self.u_prev = u
        self.u = u + self.dt * (self.nu * d2 - u * dudx + forcing)   # :468
        self.v = fft(self.u, axis=-1)                           # :469
        self.t += self.dt
        self.ioutnum += 1
        self.spec.push(self.v)
