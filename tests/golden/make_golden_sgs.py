#!/usr/bin/env python3
"""Goldens of Burger.compute_Sgs / KS.compute_Sgs (Burger.py:677-736, KS.py:385-409) recorded by RUNNING THE REAL REFERENCE
(shims of make_golden.py).  Usage: python tests/golden/make_golden_sgs.py  ->  tests/golden/sgs.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG          # noqa: E402

RB, RK = MG.RB, MG.RK

if __name__ == "__main__":
    np.seterr(over="raise", invalid="raise")
    bundle = {}
    for tag, N, nURG, forcing in [("b512", 512, 32, False), ("b256_forced", 256, 16, True), ("b1024", 1024, 64, False)]:
        b = RB.Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=24, case="turbulence", forcing=forcing, seed=42)
        b.simulate()
        b.compute_Sgs(nURG)
        p = f"{tag}/"
        bundle[p + "u0"] = b.u0.copy()
        bundle[p + "uu"] = b.uu.copy()
        bundle[p + "k"] = b.k.copy()
        bundle[p + "sgs"] = b.sgsHistory.copy()
        bundle[p + "alt"] = b.sgsHistoryAlt.copy()
        bundle[p + "alt2"] = b.sgsHistoryAlt2.copy()
        bundle[p + "randfac1"] = b.randfac1[:, :1].copy()
        bundle[p + "randfac2"] = b.randfac2[:, :1].copy()
        bundle[p + "cfg"] = np.array([N, nURG, float(forcing), 24], dtype=float)
    u0 = np.random.default_rng(5).normal(0.0, 1e-3, 256)
    pre = RK.KS(L=22.0, N=256, dt=0.25, nsteps=400, u0=u0)
    pre.simulate()
    ks = RK.KS(L=22.0, N=256, dt=0.25, nsteps=16, v0=pre.v.copy())
    ks.simulate()
    ks.fou2real()
    ks.compute_Sgs(32)
    bundle["ks256/v0"] = np.array(pre.v, dtype=np.complex128)
    bundle["ks256/uu"] = np.array(ks.uu, dtype=np.float64)
    bundle["ks256/k"] = ks.k.copy()
    bundle["ks256/sgs"] = ks.sgsHistory.copy()
    MG.save("sgs.npz", **bundle)
