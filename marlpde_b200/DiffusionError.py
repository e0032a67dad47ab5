"""``from marlpde_b200.DiffusionError import DiffusionError`` mirrors the reference module name (python/_model/DiffusionError.py)."""
from ._fd import DiffusionError  # noqa: F401
