#!/usr/bin/env python3
"""Golden vectors of the reference's DiffusionError and Laplace classes (python/_model/DiffusionError.py, Laplace.py),
recorded by RUNNING THE REAL REFERENCE in the build container (shims of make_golden.py).
Usage:  python tests/golden/make_golden_fd2.py   ->  tests/golden/fd_extra.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG          # noqa: E402
import DiffusionError as RDE      # noqa: E402
import Laplace as RLP             # noqa: E402


def diffusion_error(tag, N=32, numAgents=1, case="sinus", nsteps=40):
    d = RDE.DiffusionError(L=2 * np.pi, N=N, dt=1e-3, nu=0.05, nsteps=nsteps, case=case)
    rng = np.random.default_rng(5)
    U, A = [d.u.copy()], []
    for i in range(nsteps):
        if numAgents == 0:
            d.step()
            continue_ = True
        elif numAgents == 1:
            a = rng.uniform(-0.2, 0.2, 1)
            d.step(a.tolist(), 1)
        else:
            a = rng.uniform(-0.2, 0.2, N)
            d.step([[x] for x in a], N)
        if numAgents:
            A.append(np.atleast_1d(a).copy())
        U.append(d.u.copy())
    rw = np.atleast_1d(np.asarray(d.getMseReward(max(numAgents, 1)), dtype=np.float64)) if case == "sinus" else np.zeros(1)
    return {f"{tag}/u": np.array(U), f"{tag}/actions": np.array(A) if A else np.zeros((0, 1)), f"{tag}/reward": rw,
            f"{tag}/state": np.asarray(d.getState(1), dtype=np.float64)}


def laplace(tag, N=16, sforce="sin", ic="one", nsteps=30):
    lp = RLP.Laplace(L=2 * np.pi, N=N, dt=0.01, ic=ic, sforce=sforce, episodeLength=nsteps)
    nA = lp.N - 1
    rng = np.random.default_rng(9)
    U, A, R = [lp.u.copy()], [], []
    for i in range(nsteps):
        a = np.array([1.0, -2.0, 1.0]) / lp.dx ** 2 * (1 + 0.05 * rng.normal(size=(nA, 3)))
        lp.step(a.tolist(), nA)
        A.append(a.copy()); U.append(lp.u.copy()); R.append(np.asarray(lp.getDirectReward(nA)))
    return {f"{tag}/u": np.array(U), f"{tag}/actions": np.array(A), f"{tag}/reward": np.array(R), f"{tag}/force": lp.force.copy(),
            f"{tag}/state": np.asarray(lp.getState(nA), dtype=np.float64)}


if __name__ == "__main__":
    np.seterr(over="raise", invalid="raise")
    bundle = {}
    bundle.update(diffusion_error("de_noact", numAgents=0))
    bundle.update(diffusion_error("de_one", numAgents=1))
    bundle.update(diffusion_error("de_marl", numAgents=32))
    bundle.update(diffusion_error("de_box", numAgents=32, case="box"))
    bundle.update(laplace("lp_sin"))
    bundle.update(laplace("lp_gauss", sforce="gaussian", ic="sin"))
    MG.save("fd_extra.npz", **bundle)
