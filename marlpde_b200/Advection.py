"""``from marlpde_b200.Advection import Advection`` -- same module name as the reference's
python/_model/Advection.py.  Implementation in _fd.py."""
from ._fd import Advection  # noqa: F401
