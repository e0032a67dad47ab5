"""Pin the numpy oracle to the golden vectors recorded from the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.burger_oracle import (BurgerOracle, forcing_tables, turbulence_ic, sinus_ic, truncated_ic)
from oracle.common import grid, spectral_rel_err

STEP_CASES = {
    # name: (oracle kwargs, M, basis)
    "direct": (dict(), 32, "hat"),
    "direct_forced": (dict(forcing=True), 32, "hat"),
    "eddy": (dict(dforce=False), 32, "hat"),
    "eddy_forced": (dict(forcing=True, dforce=False), 32, "hat"),
    "eddy_forced_s4": (dict(forcing=True, dforce=False, stepper=4), 32, "hat"),
    "noact": (dict(), 0, None),
    "noact_forced": (dict(forcing=True), 0, None),
    "ssm": (dict(ssm=True), 0, None),
    "dsm": (dict(dsm=True), 0, None),
    "ssm_act": (dict(ssm=True), 32, "hat"),
    "dsm_eddy": (dict(dsm=True, dforce=False), 32, "hat"),
    "uniform8": (dict(dforce=False), 8, "uniform"),
    "hat5": (dict(), 5, "hat"),
    "one_action": (dict(dforce=False), 1, "hat"),
    "sinus64": (dict(N=64, dforce=False), 64, "hat"),
    "n16": (dict(N=16), 16, "hat"),
}


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def make(case, g, B=1):
    kw, M, basis = STEP_CASES[case]
    kw = dict(kw)
    N = kw.pop("N", 32)
    o = BurgerOracle(B=B, L=2 * np.pi, N=N, dt=1e-3, nu=0.02, **kw)
    if M:
        o.setup_basis(M, basis)
    o.set_forcing_tables(g[f"{case}/randfac1"], g[f"{case}/randfac2"])
    return o, M


@pytest.mark.parametrize("case", sorted(STEP_CASES))
def test_free_running_matches_reference(golden, case):
    g = golden("burger_steps.npz")
    o, M = make(case, g)
    U, V, F, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/Fn_old"], g[f"{case}/actions"]
    o.IC(v0=V[0][None])
    assert rel(o.Fn_old[0], F[0]) < 1e-14
    for i in range(len(U) - 1):
        o.step(A[i][None] if M else None)
        assert rel(o.v[0], V[i + 1]) < 1e-12, (case, i)
        assert rel(o.u[0], U[i + 1]) < 1e-12, (case, i)
        assert rel(o.Fn_old[0], F[i + 1]) < 1e-12, (case, i)


@pytest.mark.parametrize("case", ["eddy_forced", "dsm", "ssm_act"])
def test_teacher_forced_single_steps(golden, case):
    """One step from each reference state (Appendix C.1 of SURVEY.md)."""
    g = golden("burger_steps.npz")
    o, M = make(case, g)
    U, V, F, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/Fn_old"], g[f"{case}/actions"]
    for i in range(0, len(U) - 1, 7):
        o.IC(v0=V[i][None])
        o.u, o.Fn_old, o.ioutnum = U[i][None].copy(), F[i][None].copy(), i
        o.step(A[i][None] if M else None)
        assert rel(o.v[0], V[i + 1]) < 1e-13
        assert rel(o.u[0], U[i + 1]) < 1e-13


def test_batched_rows_equal_single(golden):
    """Row e of a batch == the same env alone, bitwise (oracle is pure broadcasting)."""
    g = golden("burger_steps.npz")
    names = ["eddy", "uniform8"]
    o1, _ = make("eddy", g)
    V, A = g["eddy/v"], g["eddy/actions"]
    ob, _ = make("eddy", g, B=3)
    v0 = np.stack([V[0], V[5], V[9]])
    ob.IC(v0=v0)
    acts = np.stack([A[0], A[1], A[2]])
    for _ in range(5):
        ob.step(acts)
    for e in range(3):
        o1.IC(v0=v0[e][None])
        for _ in range(5):
            o1.step(acts[e][None])
        assert np.array_equal(o1.v[0], ob.v[e])
        assert np.array_equal(o1.u[0], ob.u[e])


def test_initial_conditions(golden):
    g = golden("burger_steps.npz")
    x = grid(2 * np.pi, 32)
    assert rel(turbulence_ic(x, 2 * np.pi, 32, 0.0, 42), g["direct/u"][0]) < 1e-14
    x64 = grid(2 * np.pi, 64)
    assert rel(sinus_ic(x64, 2 * np.pi, 0.0), g["sinus64/u"][0]) < 1e-15
    d = golden("burger_dns.npz")
    assert rel(turbulence_ic(grid(2 * np.pi, 1024), 2 * np.pi, 1024, 0.0, 42), d["turb1024/u0"]) < 1e-14


def test_forcing_tables_match_numpy_legacy_stream(golden):
    g = golden("burger_steps.npz")
    r1, r2 = forcing_tables(42, 60)
    assert np.array_equal(r1[:, :1], g["direct_forced/randfac1"])
    assert np.array_equal(r2[:, :4], g["eddy_forced_s4/randfac2"])


def test_basis_matches_reference(golden):
    g = golden("burger_steps.npz")
    for case in ("direct", "uniform8", "hat5", "one_action", "sinus64"):
        o, M = make(case, g)
        assert np.array_equal(o.basis, g[f"{case}/basis"]), case


@pytest.mark.parametrize("ver", range(5))
@pytest.mark.parametrize("A", [1, 4, 32])
def test_states(golden, ver, A):
    g = golden("burger_states.npz")
    p = f"v{ver}_A{A}/"
    o = BurgerOracle(B=1, N=32, version=ver, numAgents=A)
    o.IC(v0=g[p + "v"][None])
    o.u, o.u_prev = g[p + "u"][None], g[p + "u_prev"][None]
    st = o.state()[0]
    ref = g[p + "state"][0] if A == 1 else g[p + "state"]      # reference wraps A==1 in a list
    assert st.shape == ref.shape
    assert rel(st, ref) < 1e-13
    o.IC(u0=g[p + "u0"][None])
    ref0 = g[p + "state0"][0] if A == 1 else g[p + "state0"]
    assert rel(o.state()[0], ref0) < 1e-13


ENV_CASES = ["spec_A1", "spec_A4", "spec_A32_v1", "spec_noise", "mse_A1", "mse_A32", "mse_noise_A4"]


@pytest.mark.parametrize("tag", ENV_CASES)
def test_environment_episode(golden, tag):
    """Re-run the recorded burger_environment.environment episode with the oracle:
    same IC hand-off, forcing tables, actions -> same states and rewards."""
    g = golden("burger_env.npz")
    p = tag + "/"
    spectral, A, noise, forcing, dforce, ver, stepper, epl, NDNS = g[p + "cfg"]
    A, ver, stepper, epl = int(A), int(ver), int(stepper), int(epl)
    o = BurgerOracle(B=1, N=32, dt=1e-3, nu=0.02, forcing=bool(forcing), dforce=bool(dforce),
                     stepper=stepper, version=ver, numAgents=A, offset=float(g[p + "offset"]))
    o.setup_basis(32, "hat")
    o.set_forcing_tables(g[p + "randfac1"], g[p + "randfac2"])
    if spectral:
        v0 = truncated_ic(g[p + "dns_v0"], g[p + "dns_k"], float(g[p + "offset"]), 32)
        assert rel(v0, g[p + "sgs_v0"]) < 1e-14
        o.IC(v0=v0[None])
    else:
        o.IC(u0=g[p + "truth_rows"][0][None])
        assert rel(o.u[0], g[p + "sgs_u0"]) < 1e-14
    assert rel(o.state()[0], g[p + "state0"]) < 1e-12
    nint = 10
    prev = 0.0
    for s in range(epl):
        a = g[p + "actions"][s][None]
        r = np.zeros((1, A))
        for _ in range(nint):
            o.step(a)
            if not spectral:
                r += o.mse_reward(g[p + "truth_rows"][o.ioutnum][None]) / nint
        st = o.state()[0]
        assert rel(st, g[p + "states"][s]) < 1e-10, (tag, s)
        if spectral:
            err = spectral_rel_err(g[p + "dns_Ek_ktt"][o.ioutnum], o.Ek_ktt_row()[0], 32)
            r = np.full((1, A), prev - err)
            prev = err
            assert rel(o.Ek_ktt_row()[0][:16], g[p + "sgs_Ek_ktt"][s + 1]) < 1e-6
        ref = np.atleast_1d(g[p + "rewards"][s])
        np.testing.assert_allclose(r[0], ref, rtol=1e-5, atol=1e-9, err_msg=f"{tag} step {s}")
    assert rel(o.u[0], g[p + "sgs_u_final"]) < 1e-10
    assert rel(o.v[0], g[p + "sgs_v_final"]) < 1e-10


def test_nested_truth_is_dns_subsampling(golden):
    """offset == 0 on nested grids: the cubic-spline truth equals DNS sub-sampling (A.8)."""
    g = golden("burger_env.npz")
    assert rel(g["mse_A1/truth_rows"], g["mse_A1/dns_uu"]) < 1e-12
