#!/usr/bin/env python3
"""1000-step free-running goldens from the REAL reference (north_star: "energy spectra after 1000 steps within 1 %",
SURVEY 8c G1 / Appendix C.2).  Same shims as make_golden.py (imported from there); build container only.

  long/burger_*  : Burger N=32 LES, turbulence IC, 3-mode forcing, eddy-viscosity actions held for 10 solver steps
                   (the environment cadence), 1000 solver steps; time-averaged spectrum Ek_ktt rows 500 and 1000, final u, v.
  long/ks_*      : KS L=22 N=64 dt=0.25 started on the attractor, direct actions held for 4 steps, 1000 steps; Ek_ktt rows.

Usage:  python tests/golden/make_golden_long.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG          # noqa: E402  (installs the shims, imports the reference classes)

RB, RK = MG.RB, MG.RK


def burger_long(tag, seed, forcing, dforce, lo, hi, bundle, nsteps=1000, hold=10, N=32, M=32):
    b = RB.Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=nsteps, case="turbulence", forcing=forcing, dforce=dforce, seed=seed)
    b.setup_basis(M, "hat")
    rng = np.random.default_rng(17)
    acts = rng.uniform(lo, hi, (nsteps // hold, M))
    u0 = b.u0.copy()
    for i in range(nsteps):
        b.step(acts[i // hold].tolist())
    b.compute_Ek()
    p = f"burger_{tag}/"
    bundle[p + "u0"] = u0
    bundle[p + "actions"] = acts
    bundle[p + "randfac1"] = b.randfac1[:, :1].copy()
    bundle[p + "randfac2"] = b.randfac2[:, :1].copy()
    bundle[p + "Ek_ktt"] = b.Ek_ktt[[250, 500, 750, 1000], :N // 2].copy()
    bundle[p + "u_final"] = b.u.copy()
    bundle[p + "v_final"] = b.v.copy()
    bundle[p + "cfg"] = np.array([seed, float(forcing), float(dforce), nsteps, hold, N, M], dtype=float)
    print(tag, "max|u| final", np.max(np.abs(b.u)), "Ek_ktt[1000,1:4]", b.Ek_ktt[1000, 1:4])


def ks_long(tag, N, M, bundle, nsteps=1000, hold=4):
    L, dt = 22.0, 0.25
    u0 = np.random.default_rng(5).normal(0.0, 1e-3, N)
    pre = RK.KS(L=L, N=N, dt=dt, nsteps=400, u0=u0)
    pre.simulate()
    v_start = pre.v.copy()
    ks = RK.KS(L=L, N=N, dt=dt, nsteps=nsteps, v0=v_start, dforce=True)
    ks.setup_basis(M, "hat")
    rng = np.random.default_rng(23)
    acts = rng.normal(0.0, 0.05, (nsteps // hold, M))
    for i in range(nsteps):
        ks.step(acts[i // hold].tolist())
    ks.compute_Ek()
    p = f"ks_{tag}/"
    bundle[p + "v0"] = v_start
    bundle[p + "actions"] = acts
    bundle[p + "Ek_ktt"] = ks.Ek_ktt[[250, 500, 750, 1000], :N // 2].copy()
    bundle[p + "v_final"] = ks.v.copy()
    bundle[p + "cfg"] = np.array([N, M, nsteps, hold], dtype=float)
    print(tag, "Ek_ktt[1000,1:4]", ks.Ek_ktt[1000, 1:4])


def ks_reward(bundle):
    """KS.getReward (KS.py:360-367) against a DNS truth: -|u - f_truth(x, t)| on the float32 row of fou2real."""
    L, dt = 22.0, 0.25
    u0 = np.random.default_rng(5).normal(0.0, 1e-3, 256)
    pre = RK.KS(L=L, N=256, dt=dt, nsteps=400, u0=u0)
    pre.simulate()
    dns = RK.KS(L=L, N=256, dt=dt, nsteps=40, v0=pre.v.copy())
    dns.simulate()
    dns.fou2real()
    g = 64
    v0 = np.concatenate((dns.vv[0, :(g + 1) // 2], dns.vv[0, -(g - 1) // 2:])) * g / 256      # ks_environment.py:52-54
    les = RK.KS(L=L, N=g, dt=dt, nsteps=40, v0=v0)
    les.setup_basis(16, "hat")
    les.setGroundTruth(dns.tt, dns.x, dns.uu)
    rng = np.random.default_rng(31)
    acts = rng.normal(0.0, 0.05, (3, 16))
    R = []
    for i in range(12):
        les.step(acts[i // 4].tolist())
        if (i + 1) % 4 == 0:
            R.append(np.array(les.getReward(), dtype=np.float64))
    bundle["ks_reward/dns_tt"], bundle["ks_reward/dns_x"] = dns.tt.copy(), dns.x.copy()
    bundle["ks_reward/dns_uu"] = np.array(dns.uu, dtype=np.float64)
    bundle["ks_reward/v0"] = np.array(v0, dtype=np.complex128)
    bundle["ks_reward/actions"] = acts
    bundle["ks_reward/rewards"] = np.array(R)
    print("ks_reward", np.array(R).shape, np.array(R)[-1, :3])


if __name__ == "__main__":
    np.seterr(over="raise", invalid="raise")
    bundle = {}
    burger_long("eddy_forced", 81, True, False, 0.02, 0.1, bundle)
    burger_long("eddy", 42, False, False, -0.01, 0.03, bundle)
    burger_long("direct_forced", 81, True, True, -0.5, 0.5, bundle)
    ks_long("n64", 64, 16, bundle)
    ks_reward(bundle)
    MG.save("long_runs.npz", **bundle)
