"""``from marlpde_b200.Laplace import Laplace`` mirrors the reference module name (python/_model/Laplace.py)."""
from ._fd import Laplace  # noqa: F401
