"""CPU oracle for the marlpde environment time-steppers.

TEST INFRASTRUCTURE ONLY.  This package is a from-scratch numpy restatement of
the arithmetic of the reference's ``python/_model`` classes (Burger / KS /
Diffusion / Advection ``step`` + forcing + state + rewards).  It exists so the
CUDA path can be checked against something that runs without a GPU and without
``/root/reference``.

Rules ("Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``marlpde_b200/`` imports it -- the product path fails
    loudly when the CUDA library is missing, it never falls back to this;
  * parity is PINNED: ``tests/golden/*.npz`` were produced by importing the
    real reference classes from ``/root/reference/python/_model`` in the build
    container (``tests/golden/make_golden.py``) and ``tests/test_oracle_*.py``
    checks every oracle function against them.

All functions are batched: fields carry a leading env axis ``[B, N]`` and each
row is computed exactly as the reference computes its single environment.
"""
