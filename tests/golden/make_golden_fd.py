#!/usr/bin/env python3
"""Golden vectors of the reference's finite-difference Burgers solver (python/_model/Burger_fd.py), recorded by RUNNING THE
REAL REFERENCE in the build container (same shims as make_golden.py, which this script imports for them).
Usage:  python tests/golden/make_golden_fd.py   ->  tests/golden/burger_fd.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG          # noqa: E402  (installs the shims, puts the reference on sys.path; generates nothing)
import Burger_fd as RF            # noqa: E402

CASES = {
    "fd_noact": dict(M=0),
    "fd_direct": dict(dforce=True, act=(-1.0, 1.0)),
    "fd_eddy": dict(dforce=False, act=(-0.01, 0.03)),
    "fd_eddy_forced": dict(dforce=False, forcing=True, act=(-0.01, 0.03)),
    "fd_ssm": dict(ssm=True, M=0),
    "fd_ssmforce": dict(dforce=True, ssmforce=True, act=(0.0, 0.3)),
    "fd_sinus64": dict(N=64, case="sinus", dforce=False, M=16, act=(-0.01, 0.03)),
    "fd_forced_s4": dict(forcing=True, stepper=4, dforce=True, act=(-0.5, 0.5)),
    "fd_dsm": dict(dsm=True, M=0),
    "fd_dsm_eddy": dict(dsm=True, dforce=False, act=(-0.01, 0.03)),
    "fd_v1": dict(dforce=False, act=(-0.01, 0.03), version=1),
    "fd_v2": dict(dforce=True, act=(-1.0, 1.0), version=2),
}


def run(N=32, nsteps=60, hold=10, case="turbulence", forcing=False, dforce=True, ssmforce=False, ssm=False, dsm=False, stepper=1,
        M=32, basis="hat", act=None, seed=42, version=0):
    b = RF.Burger_fd(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=nsteps, case=case, forcing=forcing, dforce=dforce,
                     ssmforce=ssmforce, ssm=ssm, dsm=dsm, seed=seed, s=stepper, version=version)
    if M:
        b.setup_basis(M, basis)
    rng = np.random.default_rng(11)
    U, V, A = [b.u.copy()], [np.array(b.v, dtype=np.complex128)], []
    for i in range(nsteps):
        if M and i % hold == 0:
            a = rng.uniform(act[0], act[1], M)
        if M:
            A.append(a.copy())
            b.step(a.tolist())
        else:
            b.step()
        U.append(b.u.copy()); V.append(np.array(b.v, dtype=np.complex128))
    st = np.asarray(b.getState()[0], dtype=np.float64)
    b.compute_Ek()
    return dict(u=np.array(U), v=np.array(V), actions=np.array(A) if A else np.zeros((0, max(M, 1))),
                randfac1=b.randfac1[:, :stepper].copy(), randfac2=b.randfac2[:, :stepper].copy(), state=st,
                Ek_ktt=np.asarray(b.Ek_ktt[-1], dtype=np.float64))


if __name__ == "__main__":
    np.seterr(over="raise", invalid="raise")
    bundle = {}
    for name, kw in CASES.items():
        for k, v in run(**kw).items():
            bundle[f"{name}/{k}"] = v
    MG.save("burger_fd.npz", **bundle)
