#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE REAL REFERENCE.

Runs only in the build container (needs /root/reference, read-only).  It imports the
unmodified reference classes from /root/reference/python/_model and records their
inputs/outputs to small .npz fixtures that travel with the repo; the GPU box and the
test-suite never read /root/reference.

The reference cannot run as-is on the installed SciPy (1.18): ``interp2d`` was removed
in SciPy 1.14 and the environment modules import matplotlib/korali.  Three shims are
installed BEFORE importing it (none touches the arithmetic under test):
  * ``scipy.interpolate.interp2d`` -> RectBivariateSpline-backed stand-in (SciPy's own
    documented replacement for regular grids; same FITPACK interpolating spline);
  * ``plotting`` / ``matplotlib`` -> empty stub modules;
  * ``np.random.seed(None)`` -> fixed seed, so the reference's *unseeded* draws
    (IC offset for noise>0, Burger.py:53-57) are reproducible.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import scipy.interpolate as _si

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/_model"


# --------------------------------------------------------------------------- shims
class _Interp2dShim:
    """interp2d(x, t, z[len(t), len(x)], kind) call-compatible stand-in: sorts the query
    points, returns [len(t), len(x)] squeezed to 1-D for scalar t."""

    def __init__(self, x, y, z, kind="linear"):
        k = {"linear": 1, "cubic": 3}[kind]
        self._s = _si.RectBivariateSpline(np.asarray(x), np.asarray(y), np.asarray(z).T, kx=k, ky=k, s=0)

    def __call__(self, x, y):
        x = np.sort(np.atleast_1d(np.asarray(x, dtype=float)))
        y = np.sort(np.atleast_1d(np.asarray(y, dtype=float)))
        out = self._s(x, y).T
        return out[0] if out.shape[0] == 1 else out


_si.interp2d = _Interp2dShim
for name in ("plotting", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["plotting"].makePlot = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

_seed_orig = np.random.seed
UNSEEDED = [20240229]


def _seed(s=None):
    _seed_orig(UNSEEDED[0] if s is None else s)


np.random.seed = _seed

sys.path.insert(0, REF)
import Burger as RB            # noqa: E402
import KS as RK                # noqa: E402
import Diffusion as RD         # noqa: E402
import Advection as RA         # noqa: E402
import burger_environment as BE  # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


# --------------------------------------------------------------------------- Burgers step
def burger_run(N=32, nsteps=60, hold=10, case="turbulence", forcing=False, dforce=True, ssm=False,
               dsm=False, stepper=1, M=32, basis="hat", act=None, seed=42, nu=0.02, dt=1e-3, L=2 * np.pi,
               version=0, numAgents=1, v0=None):
    kw = dict(L=L, N=N, dt=dt, nu=nu, nsteps=nsteps, forcing=forcing, dforce=dforce, ssm=ssm, dsm=dsm,
              seed=seed, s=stepper, version=version, numAgents=numAgents)
    b = RB.Burger(case=case, **kw) if v0 is None else RB.Burger(v0=v0, **kw)
    if M:
        b.setup_basis(M, basis)
    rng = np.random.default_rng(7)
    U, V, F, A = [b.u.copy()], [np.array(b.v, dtype=np.complex128)], [b.Fn_old.copy()], []
    for i in range(nsteps):
        if M and i % hold == 0:
            a = act(rng, M) if act else None
        if M and a is not None:
            A.append(a.copy())
            b.step(a.tolist() if numAgents == 1 else a.reshape(numAgents, -1).tolist())
        else:
            b.step()
        U.append(b.u.copy()); V.append(b.v.copy()); F.append(b.Fn_old.copy())
    out = dict(u=np.array(U), v=np.array(V), Fn_old=np.array(F),
               actions=np.array(A) if A else np.zeros((0, max(M, 1))),
               randfac1=b.randfac1[:, :stepper].copy(), randfac2=b.randfac2[:, :stepper].copy(),
               x=b.x, k=b.k, basis=b.basis if M else np.zeros((0, N)), tt=b.tt.copy())
    return b, out


def gen_burger_steps():
    uni = lambda lo, hi: (lambda rng, M: rng.uniform(lo, hi, M))
    cases = {
        "direct": dict(forcing=False, dforce=True, act=uni(-1, 1)),
        "direct_forced": dict(forcing=True, dforce=True, act=uni(-1, 1)),
        "eddy": dict(forcing=False, dforce=False, act=uni(-0.01, 0.03)),
        "eddy_forced": dict(forcing=True, dforce=False, act=uni(-0.01, 0.03)),
        "eddy_forced_s4": dict(forcing=True, dforce=False, stepper=4, act=uni(-0.01, 0.03)),
        "noact": dict(forcing=False, M=0),
        "noact_forced": dict(forcing=True, M=0),
        "ssm": dict(ssm=True, M=0),
        "dsm": dict(dsm=True, M=0),
        "ssm_act": dict(ssm=True, dforce=True, act=uni(-1, 1)),
        "dsm_eddy": dict(dsm=True, dforce=False, act=uni(-0.01, 0.03)),
        "uniform8": dict(dforce=False, M=8, basis="uniform", act=uni(-0.01, 0.03)),
        "hat5": dict(dforce=True, M=5, basis="hat", act=uni(-1, 1)),
        "one_action": dict(dforce=False, M=1, act=uni(0.0, 0.03)),
        "sinus64": dict(N=64, case="sinus", dforce=False, M=64, act=uni(-0.01, 0.03)),
        "n16": dict(N=16, case="sinus", dforce=True, M=16, act=uni(-1, 1)),
    }
    bundle = {}
    for name, kw in cases.items():
        _, out = burger_run(**kw)
        for k, v in out.items():
            bundle[f"{name}/{k}"] = v
    save("burger_steps.npz", **bundle)


def gen_burger_states():
    """getState for every version x numAgents (incl. the IC row where dudt == 0)."""
    bundle = {}
    for ver in range(5):
        for A in (1, 4, 32):
            b, out = burger_run(nsteps=12, version=ver, numAgents=A, dforce=False,
                                act=lambda rng, M: rng.uniform(-0.01, 0.03, M))
            st = np.array(b.getState())
            bundle[f"v{ver}_A{A}/state"] = st
            bundle[f"v{ver}_A{A}/u"] = out["u"][-1]
            bundle[f"v{ver}_A{A}/u_prev"] = out["u"][-2]
            bundle[f"v{ver}_A{A}/v"] = out["v"][-1]
            b0 = RB.Burger(L=2 * np.pi, N=32, dt=1e-3, nu=0.02, nsteps=4, case="turbulence", version=ver, numAgents=A)
            bundle[f"v{ver}_A{A}/state0"] = np.array(b0.getState())
            bundle[f"v{ver}_A{A}/u0"] = b0.u.copy()
    save("burger_states.npz", **bundle)


# --------------------------------------------------------------------------- Burgers environment
class FakeSample(dict):
    """Stand-in for the Korali sample: scripted actions, records what the env writes."""

    def __init__(self, script):
        super().__init__()
        self["Custom Settings"] = {"Mode": "Training"}
        self.script, self.i = script, 0
        self.states, self.rewards = [], []

    def update(self):
        if self.i == 0:
            self.state0 = np.array(self["State"], dtype=float)
        if self.i > 0:
            self.states.append(np.array(self["State"], dtype=float))
            self.rewards.append(np.array(self["Reward"], dtype=float))
        self["Action"] = self.script[self.i]
        self.i += 1


def gen_burger_env():
    L, T, NDNS, g, dt, nu, epl = 2 * np.pi, 0.4, 512, 32, 1e-3, 0.02, 40   # nIntermediate = 10
    bundle = {}
    made = []

    class Rec(RB.Burger):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            made.append(self)

    BE.Burger = Rec
    for tag, spectral, A, noise, forcing, dforce, ver, stepper in [
        ("spec_A1", True, 1, 0.0, True, False, 0, 1),
        ("spec_A4", True, 4, 0.0, True, False, 0, 1),
        ("spec_A32_v1", True, 32, 0.0, False, True, 1, 1),
        ("spec_noise", True, 1, 0.05, True, False, 0, 2),
        ("mse_A1", False, 1, 0.0, True, False, 0, 1),
        ("mse_A32", False, 32, 0.0, False, False, 0, 1),
        ("mse_noise_A4", False, 4, 0.05, True, False, 0, 1),
    ]:
        dns = BE.setup_dns_default(L, NDNS, T, dt, nu, "turbulence", forcing, 50, stepper)
        rng = np.random.default_rng(11)
        if dforce:
            acts = rng.uniform(-1, 1, (epl, 32))
        else:
            acts = rng.uniform(-0.01, 0.03, (epl, 32))
        script = [a.tolist() for a in acts] if A == 1 else [a.reshape(A, -1).tolist() for a in acts]
        s = FakeSample(script)
        BE.episodeCount = 0
        del made[:]
        UNSEEDED[0] = 31337
        BE.environment(s, L, T, NDNS, g, 32, dt, nu, epl, "turbulence", spectral, forcing, dforce, False,
                       noise, 50, stepper, version=ver, dns_default=[dns], numAgents=A)
        sgs = made[-1]
        assert s["Termination"] == "Terminal"
        states = s.states + [np.array(s["State"], dtype=float)]
        rewards = s.rewards + [np.array(s["Reward"], dtype=float)]
        sgs.compute_Ek()
        p = f"{tag}/"
        bundle[p + "actions"] = acts
        bundle[p + "states"] = np.array(states)
        bundle[p + "state0"] = s.state0
        bundle[p + "rewards"] = np.array(rewards)
        bundle[p + "offset"] = np.float64(sgs.offset)
        bundle[p + "sgs_v0"] = np.array(sgs.v0, dtype=np.complex128)
        bundle[p + "sgs_u0"] = np.array(sgs.u0)
        bundle[p + "sgs_u_final"] = sgs.u.copy()
        bundle[p + "sgs_v_final"] = sgs.v.copy()
        bundle[p + "sgs_Ek_ktt"] = sgs.Ek_ktt[::10, :g // 2].copy()
        bundle[p + "dns_Ek_ktt"] = dns.Ek_ktt[:, :g // 2].copy()
        bundle[p + "dns_v0"] = np.array(dns.v0, dtype=np.complex128)
        bundle[p + "dns_k"] = dns.k.copy()
        bundle[p + "randfac1"] = dns.randfac1[:, :stepper].copy()
        bundle[p + "randfac2"] = dns.randfac2[:, :stepper].copy()
        bundle[p + "cfg"] = np.array([spectral, A, noise, forcing, dforce, ver, stepper, epl, NDNS], dtype=float)
        if not spectral:
            # the truth the LES was scored against at every sub-step (Burger.py:581-588)
            newx = sgs.x + sgs.offset
            newx[newx > L] -= L
            newx[newx < 0] += L
            order = np.argsort(np.argsort(newx))            # interp2d sorts its x; undo per piece
            midx = np.argmax(newx)
            rows = []
            for t in sgs.tt:
                if midx == len(newx) - 1:
                    rows.append(sgs.f_truth(newx, t))
                else:
                    rows.append(np.concatenate((sgs.f_truth(newx[:midx + 1], t), sgs.f_truth(newx[midx + 1:], t))))
            bundle[p + "truth_rows"] = np.array(rows)
            bundle[p + "dns_uu"] = dns.uu[:, ::NDNS // g].copy() if noise == 0 else dns.uu[::10].copy()
            bundle[p + "dns_tt"] = dns.tt.copy()
            bundle[p + "dns_x"] = dns.x.copy()
    BE.Burger = RB.Burger
    save("burger_env.npz", **bundle)


# --------------------------------------------------------------------------- DNS
def gen_burger_dns():
    bundle = {}
    for tag, kw in {
        "turb1024": dict(N=1024, case="turbulence", nsteps=200),
        "sinus512": dict(N=512, case="sinus", nsteps=200),
        "turb256_forced": dict(N=256, case="turbulence", nsteps=200, forcing=True),
        "forced_L100": dict(N=256, L=100.0, dt=0.01, case="forced", nsteps=120, forcing=True, s=20),
        "turb128": dict(N=128, case="turbulence", nsteps=100),
        "turb2048": dict(N=2048, case="turbulence", nsteps=40),
    }.items():
        kw.setdefault("L", 2 * np.pi); kw.setdefault("dt", 1e-3)
        b = RB.Burger(nu=0.02, seed=42, **kw)
        u0, v0 = b.u0.copy(), np.array(b.v0)
        b.simulate()
        b.compute_Ek()
        p = tag + "/"
        st = b.stepper
        bundle[p + "u0"] = u0
        bundle[p + "rows"] = b.uu[::20].copy()
        bundle[p + "v_final"] = b.v.copy()
        bundle[p + "Fn_old_final"] = b.Fn_old.copy()
        bundle[p + "Ek_ktt"] = b.Ek_ktt[::20, :64].copy()
        bundle[p + "Ek_tt"] = b.Ek_tt[::20].copy()
        bundle[p + "randfac1"] = b.randfac1[:, :st].copy()
        bundle[p + "randfac2"] = b.randfac2[:, :st].copy()
        bundle[p + "cfg"] = np.array([kw["N"], kw["L"], kw["dt"], kw["nsteps"], float(kw.get("forcing", False)), st])
    save("burger_dns.npz", **bundle)


# --------------------------------------------------------------------------- KS
def gen_ks():
    bundle = {}
    for tag, N, M, dforce, nrec in [("n64", 64, 16, True, 80), ("n32", 32, 32, True, 80),
                                     ("n64_noact", 64, 0, True, 80), ("n256", 256, 0, True, 40),
                                     ("n64_eddy", 64, 64, False, 40), ("n1024", 1024, 0, True, 20)]:
        L, dt = 22.0, 0.25
        u0 = np.random.default_rng(5).normal(0.0, 1e-3, N)
        pre = RK.KS(L=L, N=N, dt=dt, nsteps=400, u0=u0)
        pre.simulate()                                          # reach the attractor
        v_start = pre.v.copy()
        ks = RK.KS(L=L, N=N, dt=dt, nsteps=nrec, v0=v_start, dforce=dforce)
        if M:
            ks.setup_basis(M, "hat")
        rng = np.random.default_rng(9)
        V, S, A, E = [ks.v.copy()], [ks.getState().copy()], [], []
        for i in range(nrec):
            if M:
                if i % 4 == 0:
                    a = rng.normal(0.0, 0.05, M)
                A.append(a.copy())
                ks.step(a.tolist())
            else:
                ks.step()
            V.append(ks.v.copy())
            if (i + 1) % 4 == 0:                                # env cadence: compute_Ek + getState
                ks.compute_Ek()
                E.append(ks.Ek_ktt[ks.ioutnum, :N // 2].copy())
                S.append(ks.getState().copy())
        p = tag + "/"
        bundle[p + "v"] = np.array(V)
        bundle[p + "states"] = np.array(S)
        bundle[p + "actions"] = np.array(A) if A else np.zeros((0, 1))
        bundle[p + "Ek_ktt"] = np.array(E)
        for name in ("E", "E2", "Q", "f1", "f2", "f3"):
            bundle[p + name] = getattr(ks, name)
        bundle[p + "g"] = ks.g
        bundle[p + "cfg"] = np.array([N, L, dt, M, float(dforce), nrec])
    save("ks.npz", **bundle)


# --------------------------------------------------------------------------- FD
def gen_fd():
    bundle = {}
    rng = np.random.default_rng(3)
    N, L, dt, nu = 32, 2 * np.pi, 0.01, 0.1
    # Diffusion
    for tag, case, mode, A in [("plain", "sinus", None, 1), ("lap", "sinus", "global", 1),
                               ("point_A1", "box", "point", 1), ("point_A4", "gaussian", "point", 4),
                               ("point_AN", "sinus", "point", 32), ("implicit", "box", "implicit", 1)]:
        d = RD.Diffusion(L=L, N=N, dt=dt, nu=nu, nsteps=30, case=case, implicit=(mode == "implicit"))
        U, acts, R, D, S = [d.u.copy()], [], [], [], [np.array(d.getState(A))]
        for i in range(30):
            if mode in (None, "implicit"):
                d.step()
            elif mode == "global":
                a = np.array([-2.0 + 0.1 * rng.normal()])
                acts.append(a); d.step(a.tolist())
            else:
                a = -2.0 + 0.2 * rng.normal(size=N)
                acts.append(a)
                d.step(a.tolist() if A == 1 else a.reshape(A, -1).tolist(), numAgents=A)
            U.append(d.u.copy())
            S.append(np.array(d.getState(A)))
            if case == "sinus":
                R.append(np.atleast_1d(np.array(d.getMseReward(A))))
            if A == N:
                D.append(np.array(d.getDirectReward(A)))
        p = f"diff_{tag}/"
        bundle[p + "u"] = np.array(U); bundle[p + "actions"] = np.array(acts) if acts else np.zeros((0, 1))
        bundle[p + "states"] = np.array(S)
        bundle[p + "mse"] = np.array(R) if R else np.zeros((0, 1))
        bundle[p + "direct"] = np.array(D) if D else np.zeros((0, 1))
        bundle[p + "solution"] = d.solution.copy()
        bundle[p + "cfg"] = np.array([N, L, dt, nu, A])
    # Advection
    dt, nu = 0.01, 1.0
    for tag, mode, A in [("lax", None, 1), ("global", "global", 1), ("point_A1", "point", 1), ("point_A4", "point", 4)]:
        a_ = RA.Advection(L=L, N=N, dt=dt, nu=nu, nsteps=30, case="sinus")
        al = a_.alpha
        U, acts, R, S = [a_.u.copy()], [], [], [np.array(a_.getState(A))]
        for i in range(30):
            if mode is None:
                a_.step()
            elif mode == "global":
                a = np.array([0.5 + 0.5 * al, 0.5 - 0.5 * al]) + 0.01 * rng.normal(size=2)
                acts.append(a); a_.step(a.tolist())
            else:
                a = np.tile([0.5 - 0.5 * al, 0.5 + 0.5 * al], N) + 0.01 * rng.normal(size=2 * N)
                acts.append(a)
                a_.step(a.tolist() if A == 1 else a.reshape(A, -1).tolist(), numAgents=A)
            U.append(a_.u.copy()); S.append(np.array(a_.getState(A)))
            R.append(np.atleast_1d(np.array(a_.getMseReward(A))))
        p = f"adv_{tag}/"
        bundle[p + "u"] = np.array(U); bundle[p + "actions"] = np.array(acts) if acts else np.zeros((0, 1))
        bundle[p + "states"] = np.array(S); bundle[p + "mse"] = np.array(R)
        bundle[p + "solution"] = a_.solution.copy()
        bundle[p + "cfg"] = np.array([N, L, dt, nu, A])
    save("fd.npz", **bundle)


# --------------------------------------------------------------------------- KS environment
def gen_ks_env():
    """The reference ks_environment.setup_dns_default + environment with a fake Korali sample
    (DNS N = 256 instead of the script default to keep the fixture small)."""
    import ks_environment as KE
    bundle = {}
    UNSEEDED[0] = 777
    NDNS, dt, g, epl = 256, 0.25, 32, 50          # nIntermediate = int(500/0.25/50) = 40
    captured = {}
    orig_ic = RK.KS.IC

    def spy_ic(self, u0=None, v0=None, case='noise', seed=42):
        r = orig_ic(self, u0=u0, v0=v0, case=case, seed=seed)
        if u0 is None and v0 is None and 'transient_u0' not in captured:
            captured['transient_u0'] = self.u0.copy()
        return r

    RK.KS.IC = spy_ic
    KE.KS = RK.KS
    dns = KE.setup_dns_default(NDNS, dt, 1.0, 42)
    RK.KS.IC = orig_ic
    for tag, dforce in (("direct", True), ("eddy", False)):
        rng = np.random.default_rng(21)
        acts = rng.normal(0.0, 0.05 if dforce else 0.01, (epl, 16))
        s = FakeSample([a.tolist() for a in acts])
        KE.environment(s, NDNS, g, 16, dt, 1.0, epl, dforce, 42, dns)
        assert s["Termination"] == "Terminal"
        p = f"{tag}/"
        bundle[p + "actions"] = acts
        bundle[p + "state0"] = s.state0
        bundle[p + "states"] = np.array(s.states + [np.array(s["State"], dtype=float)])
        bundle[p + "rewards"] = np.array(s.rewards + [np.array(s["Reward"], dtype=float)])
    bundle["transient_u0"] = captured['transient_u0']
    bundle["dns_u0"] = np.array(dns.u0)
    bundle["dns_vv0"] = dns.vv[0].copy()
    bundle["dns_Ek_ktt"] = dns.Ek_ktt[:, :g // 2].copy()
    bundle["dns_uu_last"] = np.array(dns.uu[-1])
    bundle["cfg"] = np.array([NDNS, dt, g, epl, 16])
    save("ks_env.npz", **bundle)


if __name__ == "__main__":
    np.seterr(over="raise", invalid="raise")
    gen_burger_steps()
    gen_burger_states()
    gen_burger_env()
    gen_burger_dns()
    gen_ks()
    gen_fd()
    gen_ks_env()
