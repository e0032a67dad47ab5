"""Pin the KS / Diffusion / Advection / DNS oracles to golden vectors recorded from the
real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.burger_oracle import BurgerOracle
from oracle.ks_oracle import KSOracle, etdrk4_tables
from oracle.fd_oracle import DiffusionOracle, AdvectionOracle


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


KS_CASES = ["n64", "n32", "n64_noact", "n256", "n64_eddy", "n1024"]


@pytest.mark.parametrize("tag", KS_CASES)
def test_ks_tables(golden, tag):
    g = golden("ks.npz")
    N, L, dt, M, dforce, nrec = g[tag + "/cfg"]
    T = etdrk4_tables(L, int(N), dt)
    for name in ("E", "E2", "Q", "f1", "f2", "f3", "g"):
        assert np.array_equal(T[name], g[f"{tag}/{name}"]), name


@pytest.mark.parametrize("tag", KS_CASES)
def test_ks_teacher_forced_and_state(golden, tag):
    """KS is chaotic: compare one step from each reference state (1e-10 on v), the
    float32 state (Q7) and the float32 spectrum chain at the environment cadence."""
    g = golden("ks.npz")
    N, L, dt, M, dforce, nrec = g[tag + "/cfg"]
    N, M, nrec = int(N), int(M), int(nrec)
    V, A, S, E = g[tag + "/v"], g[tag + "/actions"], g[tag + "/states"], g[tag + "/Ek_ktt"]
    o = KSOracle(B=1, L=L, N=N, dt=dt, dforce=bool(dforce))
    if M:
        o.setup_basis(M, "hat")
    o.IC(v0=V[0][None])

    def state_ok(st, ref):
        # the reference state is a second difference of a FLOAT32 field (Q7): its own rounding
        # noise is ~ulp32(max|u|)/dx^2, which is the tightest meaningful pin
        noise = 8 * np.finfo(np.float32).eps * np.max(np.abs(o.uu_row)) / o.dx ** 2
        return np.max(np.abs(st - ref)) <= noise + 2e-6 * np.max(np.abs(ref))

    assert state_ok(o.state()[0], S[0])
    for i in range(nrec):
        o.v = V[i][None].copy()   # re-anchor on the reference state; counters/spectrum keep running
        o.step(A[i][None] if M else None)
        # dforce=False multiplies by a stencil of the FLOAT32 uu row (KS.py:241-245): a 1-ulp(f32)
        # flip in that row moves v by ~1e-10, so that case is pinned at 1e-8
        assert rel(o.v[0], V[i + 1]) < (1e-10 if dforce else 1e-8), (tag, i)
        if (i + 1) % 4 == 0:
            j = (i + 1) // 4
            np.testing.assert_allclose(o.Ek_ktt_row()[0][:N // 2], E[j - 1], rtol=2e-5, atol=1e-30)
            st = o.state()[0]
            assert state_ok(st, S[j]), (tag, i)


def test_ks_free_running_short(golden):
    g = golden("ks.npz")
    V, A = g["n64/v"], g["n64/actions"]
    o = KSOracle(B=1, L=22.0, N=64, dt=0.25)
    o.setup_basis(16, "hat")
    o.IC(v0=V[0][None])
    for i in range(20):
        o.step(A[i][None])
    assert rel(o.v[0], V[20]) < 1e-9


DNS_CASES = ["turb1024", "sinus512", "turb256_forced", "forced_L100", "turb128", "turb2048"]


@pytest.mark.parametrize("tag", DNS_CASES)
def test_burgers_dns(golden, tag):
    g = golden("burger_dns.npz")
    N, L, dt, nsteps, forcing, st = g[tag + "/cfg"]
    N, nsteps, st = int(N), int(nsteps), int(st)
    o = BurgerOracle(B=1, L=L, N=N, dt=dt, nu=0.02, forcing=bool(forcing), stepper=st)
    o.set_forcing_tables(g[tag + "/randfac1"], g[tag + "/randfac2"])
    o.IC(u0=g[tag + "/u0"][None])
    rows, Ek = g[tag + "/rows"], g[tag + "/Ek_ktt"]
    for i in range(1, nsteps + 1):
        o.step()
        if i % 20 == 0:
            assert rel(o.u[0], rows[i // 20]) < 1e-11, (tag, i)
            np.testing.assert_allclose(o.Ek_ktt_row()[0][:64], Ek[i // 20], rtol=1e-5, atol=1e-30)
    assert rel(o.v[0], g[tag + "/v_final"]) < 1e-11
    assert rel(o.Fn_old[0], g[tag + "/Fn_old_final"]) < 1e-11


DIFF = {"plain": 1, "lap": 1, "point_A1": 1, "point_A4": 4, "point_AN": 32, "implicit": 1}


@pytest.mark.parametrize("tag", sorted(DIFF))
def test_diffusion(golden, tag):
    g = golden("fd.npz")
    p = f"diff_{tag}/"
    N, L, dt, nu, A = g[p + "cfg"]
    N, A = int(N), int(A)
    U, acts, S = g[p + "u"], g[p + "actions"], g[p + "states"]
    o = DiffusionOracle(B=1, L=L, N=N, dt=dt, nu=nu, implicit=(tag == "implicit"))
    o.IC(U[0][None])
    for i in range(len(U) - 1):
        o.step(acts[i][None] if len(acts) else None)
        assert rel(o.u[0], U[i + 1]) < 1e-13, (tag, i)
        st = o.state(A)[0]
        assert rel(st, S[i + 1].reshape(st.shape)) < 1e-13
        if len(g[p + "mse"]):
            np.testing.assert_allclose(o.mse_reward(o.analytic(), A)[0], g[p + "mse"][i], rtol=1e-9, atol=1e-30)
            assert rel(o.analytic()[0], g[p + "solution"][i + 1]) < 1e-14
        if len(g[p + "direct"]):
            np.testing.assert_allclose(o.direct_reward()[0], g[p + "direct"][i], rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("tag", ["lax", "global", "point_A1", "point_A4"])
def test_advection(golden, tag):
    g = golden("fd.npz")
    p = f"adv_{tag}/"
    N, L, dt, nu, A = g[p + "cfg"]
    N, A = int(N), int(A)
    U, acts, S = g[p + "u"], g[p + "actions"], g[p + "states"]
    o = AdvectionOracle(B=1, L=L, N=N, dt=dt, nu=nu)
    o.IC(U[0][None])
    for i in range(len(U) - 1):
        o.step(acts[i][None] if len(acts) else None)
        assert rel(o.u[0], U[i + 1]) < 1e-13, (tag, i)
        st = o.state(A)[0]
        assert rel(st, S[i + 1].reshape(st.shape)) < 1e-13
        np.testing.assert_allclose(o.mse_reward(A)[0], g[p + "mse"][i], rtol=1e-9, atol=1e-30)
        assert rel(o.analytic(), g[p + "solution"][i + 1]) < 1e-14


def test_known_answers():
    """action -2 == standard Laplacian; Lax weights == FDstep; both
    converge to the analytic sinus solutions."""
    N, L = 64, 2 * np.pi
    x = np.linspace(0, L, N, endpoint=False)
    d0 = DiffusionOracle(N=N, L=L, dt=1e-3, nu=0.1); d0.IC(np.sin(x)[None])
    d1 = DiffusionOracle(N=N, L=L, dt=1e-3, nu=0.1); d1.IC(np.sin(x)[None])
    for _ in range(200):
        d0.step(); d1.step(np.array([[-2.0]]))
    assert rel(d0.u, d1.u) < 1e-13
    assert rel(d0.u, d0.analytic()) < 1e-3
    a0 = AdvectionOracle(N=N, L=L, dt=0.09, nu=1.0); a0.IC(np.sin(x)[None])     # Courant ~0.92
    a1 = AdvectionOracle(N=N, L=L, dt=0.09, nu=1.0); a1.IC(np.sin(x)[None])
    al = a0.alpha
    for _ in range(20):
        a0.step(); a1.step(np.array([[0.5 + 0.5 * al, 0.5 - 0.5 * al]]))
    assert rel(a0.u, a1.u) < 1e-13
    assert rel(a0.u[0], a0.analytic()) < 2e-2
