"""marlpde_b200 -- B200-native batched environment time-steppers with the class API of
wadaniel/marlpde's ``python/_model`` (Burger, Burger_fd, KS, Diffusion, Advection, DiffusionError, Laplace).

The arithmetic runs in ``libmarlpde_b200.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/marlpde_b200.h``); importing the package does not need a GPU, constructing an
environment does.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .Burger import Burger  # noqa: F401
from .Burger_fd import Burger_fd  # noqa: F401
from .KS import KS  # noqa: F401
from ._fd import Diffusion, Advection, DiffusionError, Laplace  # noqa: F401

__all__ = ["Burger", "Burger_fd", "KS", "Diffusion", "Advection", "DiffusionError", "Laplace"]
