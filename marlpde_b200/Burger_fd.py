"""Batched finite-difference Burgers solver: the reference's ``Burger_fd`` class
(/root/reference/python/_model/Burger_fd.py) on the GPU.  Same constructor keywords, basis, IC, state, reward and
spectrum methods as ``Burger`` (the reference classes differ only in ``step``: explicit Euler in time, first-order
upwind / centred differences in space, v = fft(u) refreshed every step, Burger_fd.py:335-476, and in ``ssmforce``
actually being applied, :447-455).  N <= 256 (warp-resident kernels); state versions 0, 1, 2 as in the reference
(Burger_fd.py:590-640); static / dynamic Smagorinsky closures, forcing and every action mode.  ``nunoise=True`` draws nu
from U(0.015, 0.025) per environment out of an UNSEEDED generator, exactly like the reference (Burger_fd.py:53,81-82:
``np.random.seed(None)`` precedes it) -- pass ``nu=`` per environment to pin it."""
import numpy as np
import torch

from . import _lib as LB
from .Burger import Burger


class Burger_fd(Burger):
    def __init__(self, *args, **kw):
        if kw.get("version", 0) not in (0, 1, 2):
            raise SystemExit("[Burger_fd] Version not recognized")          # Burger_fd.py:642-644
        self._ssmforce_flag = bool(kw.get("ssmforce", False))
        nunoise = bool(kw.pop("nunoise", False))
        super().__init__(*args, **kw)
        if nunoise:                                                          # Burger_fd.py:81-82 (unseeded draw)
            self.set_nu(np.random.default_rng().uniform(0.015, 0.025, self.nenvs))

    def set_nu(self, nu):
        """Per-environment viscosity (what ``nunoise`` draws; pass it here to reproduce a run)."""
        self._nu = np.broadcast_to(np.asarray(nu, dtype=np.float64), (self.nenvs,)).copy()
        LB.check(self._lib.mpde_set_nu(self._h, LB.as_dp(np.ascontiguousarray(self._nu)), self.nenvs))

    def _extra_flags(self):
        return LB.FD | (LB.SSMFORCE if self._ssmforce_flag else 0)

    @property
    def u(self):
        """The field itself is the primary variable of the finite-difference solver (kept in the u_prev slot)."""
        return self._squeeze(self._get(LB.FIELD_U_PREV, (self.nenvs, self.N), self.dtype))
