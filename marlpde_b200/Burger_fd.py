"""Batched finite-difference Burgers solver: the reference's ``Burger_fd`` class
(/root/reference/python/_model/Burger_fd.py) on the GPU.  Same constructor keywords, basis, IC, state, reward and
spectrum methods as ``Burger`` (the reference classes differ only in ``step``: explicit Euler in time, first-order
upwind / centred differences in space, v = fft(u) refreshed every step, Burger_fd.py:335-476, and in ``ssmforce``
actually being applied, :447-455).  Supported here: N <= 256, state versions 0 and 2, every closure / forcing /
action mode except the dynamic Smagorinsky model and ``nunoise`` (which draws from the unseeded generator)."""
import torch

from . import _lib as LB
from .Burger import Burger


class Burger_fd(Burger):
    def __init__(self, *args, **kw):
        if kw.get("dsm"):
            raise NotImplementedError("Burger_fd(dsm=True) is not available on the GPU path")
        if kw.get("nunoise"):
            raise NotImplementedError("Burger_fd(nunoise=True) draws nu from the unseeded generator; pass nu explicitly")
        if kw.get("version", 0) not in (0, 2):
            raise NotImplementedError("Burger_fd on the GPU path supports state versions 0 and 2")
        self._ssmforce_flag = bool(kw.get("ssmforce", False))
        super().__init__(*args, **kw)

    def _extra_flags(self):
        return LB.FD | (LB.SSMFORCE if self._ssmforce_flag else 0)

    @property
    def u(self):
        """The field itself is the primary variable of the finite-difference solver (kept in the u_prev slot)."""
        return self._squeeze(self._get(LB.FIELD_U_PREV, (self.nenvs, self.N), self.dtype))
