"""``from marlpde_b200.Diffusion import Diffusion`` -- same module name as the reference's
python/_model/Diffusion.py.  Implementation in _fd.py."""
from ._fd import Diffusion  # noqa: F401
