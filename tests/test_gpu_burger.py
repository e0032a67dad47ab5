"""GPU parity tests of the batched Burgers stepper: CUDA path (through the C ABI) against
(1) the golden vectors recorded from the real reference and (2) the numpy oracle on the
same seeded inputs.  Tolerances: 1e-10 relative on u, v, Fn_old, state (fp64 path);
1e-5 / 1e-6 on the float32 spectrum chain and rewards (reference quirk Q6)."""
import numpy as np
import pytest
import torch

from tests.test_oracle_burger import STEP_CASES

pytestmark = pytest.mark.gpu

TWO_PI = 2 * np.pi


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def make_env(case, g, B=1, nsteps=60, **extra):
    from marlpde_b200 import Burger
    kw, M, basis = STEP_CASES[case]
    kw = dict(kw)
    N = kw.pop("N", 32)
    stepper = kw.pop("stepper", 1)
    env = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=nsteps, case="zero", s=stepper, nenvs=B, **kw, **extra)
    if M:
        env.setup_basis(M, basis)
    if kw.get("forcing"):
        env.randfac1 = g[f"{case}/randfac1"]
        env.randfac2 = g[f"{case}/randfac2"]
    return env, M


@pytest.mark.parametrize("case", sorted(STEP_CASES))
def test_free_running_vs_reference(golden, case):
    """60 steps from the reference IC with the reference's actions: u, v, Fn_old at every step."""
    g = golden("burger_steps.npz")
    env, M = make_env(case, g)
    U, V, F, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/Fn_old"], g[f"{case}/actions"]
    env.IC(v0=V[0])
    assert rel(env.Fn_old, F[0]) < 1e-13
    assert rel(env.u, U[0]) < 1e-13
    worst = 0.0
    for i in range(len(U) - 1):
        env.step(A[i] if M else None)
        if i % 6 == 5 or i < 3:
            worst = max(worst, rel(env.v, V[i + 1]), rel(env.u, U[i + 1]), rel(env.Fn_old, F[i + 1]))
    assert worst < 1e-10, (case, worst)
    assert int(env.status) == 0 if env.nenvs == 1 else True


@pytest.mark.parametrize("case", ["eddy_forced", "direct", "dsm", "ssm_act", "sinus64", "n16"])
def test_teacher_forced_single_steps(golden, case):
    """One GPU step from each reference state, all states as one batch."""
    from marlpde_b200 import _lib as LB
    g = golden("burger_steps.npz")
    U, V, F, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/Fn_old"], g[f"{case}/actions"]
    kw = STEP_CASES[case][0]
    if kw.get("stepper", 1) != 1:
        pytest.skip("column index differs per row")
    idx = np.arange(0, len(U) - 1, 1)
    B = len(idx)
    env, M = make_env(case, g, B=B)
    env.IC(v0=V[idx])
    env._set(LB.FIELD_FN_OLD, torch.view_as_real(torch.as_tensor(F[idx], device=env.device).contiguous()))
    env.step(A[idx] if M else None)
    assert rel(env.v, V[idx + 1]) < 1e-10
    assert rel(env.u, U[idx + 1]) < 1e-10
    assert rel(env.Fn_old, F[idx + 1]) < 1e-10


def test_fused_substeps_equal_single_steps(golden):
    """step_n(a, 10) == 10 x step(a), bitwise (same kernel arithmetic, state kept in registers)."""
    g = golden("burger_steps.npz")
    V, A = g["eddy_forced/v"], g["eddy_forced/actions"]
    e1, _ = make_env("eddy_forced", g, B=4)
    e2, _ = make_env("eddy_forced", g, B=4)
    v0 = V[[0, 7, 19, 33]]
    a = A[[0, 1, 2, 3]]
    e1.IC(v0=v0); e2.IC(v0=v0)
    e1.step_n(a, 10)
    for _ in range(10):
        e2.step(a)
    assert torch.equal(e1.v, e2.v)
    assert torch.equal(e1.u, e2.u)
    assert torch.equal(e1.Fn_old, e2.Fn_old)


@pytest.mark.parametrize("B", [1, 2, 3, 5, 64, 257])
def test_batch_invariance_bitwise(golden, B):
    """Env e of a batch == the same env alone, bitwise, whatever the pairing/packing."""
    g = golden("burger_steps.npz")
    V, A = g["eddy_forced/v"], g["eddy_forced/actions"]
    rng = np.random.default_rng(B)
    rows = rng.integers(0, len(V), B)
    acts = A[rng.integers(0, len(A), B)]
    eb, _ = make_env("eddy_forced", g, B=B)
    eb.IC(v0=V[rows])
    eb.step_n(acts, 5)
    vb, ub = eb.v.reshape(B, -1), eb.u.reshape(B, -1)
    for e in sorted(set([0, B // 2, B - 1])):
        e1, _ = make_env("eddy_forced", g, B=1)
        e1.IC(v0=V[rows[e]])
        e1.step_n(acts[e], 5)
        assert torch.equal(e1.v, vb[e]), (B, e)
        assert torch.equal(e1.u, ub[e]), (B, e)


@pytest.mark.parametrize("ver", range(5))
@pytest.mark.parametrize("A", [1, 4, 32])
def test_states(golden, ver, A):
    from marlpde_b200 import Burger, _lib as LB
    g = golden("burger_states.npz")
    p = f"v{ver}_A{A}/"
    env = Burger(L=TWO_PI, N=32, dt=1e-3, nu=0.02, nsteps=20, case="zero", version=ver, numAgents=A)
    env.IC(v0=g[p + "v"])
    env._set(LB.FIELD_U_PREV, torch.as_tensor(g[p + "u_prev"][None], device=env.device).contiguous())
    env._set(LB.FIELD_IOUTNUM, torch.ones(1, dtype=torch.int32, device=env.device))
    env.ioutnum = 1
    st = np.array(env.getState())
    assert st.shape == g[p + "state"].shape
    assert rel(st, g[p + "state"]) < 1e-10
    env.IC(u0=g[p + "u0"])
    st0 = np.array(env.getState())
    assert rel(st0, g[p + "state0"]) < 1e-10


ENV_CASES = ["spec_A1", "spec_A4", "spec_A32_v1", "spec_noise", "mse_A1", "mse_A32", "mse_noise_A4"]


@pytest.mark.parametrize("history", [True, False])
@pytest.mark.parametrize("tag", ENV_CASES)
def test_environment_episode(golden, tag, history):
    """The recorded burger_environment.environment episode (IC hand-off, forcing tables, scripted
    actions, nIntermediate = 10): states and rewards of every RL step, one launch per RL step.  With history=False the
    training specialisations of the kernel run (LEAN / HOT for the spectral reward, the multi-agent MSE variant for the
    truth-table cases) instead of the generic one."""
    from marlpde_b200 import Burger
    from oracle.burger_oracle import truncated_ic
    g = golden("burger_env.npz")
    p = tag + "/"
    spectral, A, noise, forcing, dforce, ver, stepper, epl, NDNS = g[p + "cfg"]
    A, ver, stepper, epl = int(A), int(ver), int(stepper), int(epl)
    off = float(g[p + "offset"])
    env = Burger(L=TWO_PI, N=32, dt=1e-3, nu=0.02, tend=0.4, case="zero", forcing=bool(forcing), dforce=bool(dforce),
                 s=stepper, version=ver, numAgents=A, offset=off, history=history)
    env.setup_basis(32, "hat")
    env.randfac1, env.randfac2 = g[p + "randfac1"], g[p + "randfac2"]
    if spectral:
        env.IC(v0=truncated_ic(g[p + "dns_v0"], g[p + "dns_k"], off, 32))
        env.set_spectrum_reference(g[p + "dns_Ek_ktt"])
    else:
        env.IC(u0=g[p + "truth_rows"][0])
        env.set_truth_table(g[p + "truth_rows"])
    st0 = np.array(env.getState())
    assert rel(st0.reshape(-1), g[p + "state0"].reshape(-1)) < 1e-10
    for s in range(epl):
        state, reward = env.step_n(g[p + "actions"][s], 10)
        assert rel(state.reshape(-1), g[p + "states"][s].reshape(-1)) < 1e-10, (tag, s)
        ref = np.atleast_1d(g[p + "rewards"][s])
        np.testing.assert_allclose(reward[0].cpu().numpy(), ref, rtol=1e-5, atol=1e-9, err_msg=f"{tag} step {s}")
        if spectral:
            assert rel(env.Ek_ktt_row()[:16], g[p + "sgs_Ek_ktt"][s + 1]) < 1e-6
    assert rel(env.u, g[p + "sgs_u_final"]) < 1e-10
    assert rel(env.v, g[p + "sgs_v_final"]) < 1e-10
    assert env.ioutnum == epl * 10
    if history:          # history rows written by the kernel
        assert rel(env.uu[env.ioutnum], g[p + "sgs_u_final"]) < 1e-10


def test_wavenumber_table_is_numpy_fftfreq():
    from marlpde_b200 import Burger, _lib as LB
    for N, L_ in [(32, TWO_PI), (64, 100.0), (128, 22.0)]:
        env = Burger(L=L_, N=N, nsteps=2, case="zero")
        k = env._get(LB.FIELD_K, (N,), torch.float64).cpu().numpy()
        assert np.array_equal(k, env.k), (N, L_)


def test_blow_up_is_truncated_not_fatal(golden):
    """A diverging env reports status TRUNCATED / inf state; its pair partner is unaffected."""
    g = golden("burger_steps.npz")
    V = g["direct/v"]
    env, M = make_env("direct", g, B=2)
    env.IC(v0=V[[0, 0]])
    acts = np.zeros((2, 32)); acts[1] = 1e200
    st, _ = env.step_n(acts, 5)
    status = env.status.cpu().numpy()
    assert status.tolist() == [0, 1]
    assert torch.isinf(st[1]).all() and torch.isfinite(st[0]).all()
    alone, _ = make_env("direct", g, B=1)
    alone.IC(v0=V[0]); alone.step_n(acts[0], 5)
    assert torch.equal(alone.v, env.v[0])


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (B = 4096, N = 32, forcing, spectral reward, nIntermediate 10):
    size-independent properties -- batch rows equal a small re-run bitwise, energy is finite,
    identical envs give identical results, Hermitian symmetry of v, reward sums telescope."""
    from marlpde_b200 import Burger
    B, N = 4096, 32
    rng = np.random.default_rng(0)
    seeds = 42 + (np.arange(B) % 7)
    env = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, tend=5, case="turbulence", forcing=True, dforce=False, seed=seeds,
                 nenvs=B, history=False)
    env.setup_basis(32, "hat")
    ref = np.abs(rng.normal(1.0, 0.1, (5001, 16))) * 1e-3 + 1e-6
    env.set_spectrum_reference(ref)
    acts = rng.uniform(-0.005, 0.02, (B, 32))
    tot = torch.zeros(B, 1, dtype=torch.float64, device=env.device)
    for _ in range(3):
        st, rw = env.step_n(acts, 10)
        tot += rw
    assert torch.isfinite(st).all() and torch.isfinite(rw).all()
    v = env.v
    assert torch.equal(v[:, 1:16], v[:, 17:].flip(1).conj())          # exact Hermitian symmetry
    # envs with the same seed and the same actions are bit-identical
    acts2 = np.tile(acts[:7], (B // 7 + 1, 1))[:B]
    env2 = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, tend=5, case="turbulence", forcing=True, dforce=False, seed=seeds,
                  nenvs=B, history=False)
    env2.setup_basis(32, "hat")
    env2.set_spectrum_reference(ref)
    env2.step_n(acts2, 10)
    assert torch.equal(env2.v[:7], env2.v[7:14]) and torch.equal(env2.v[:7], env2.v[B - B % 7 - 7:B - B % 7])
    # telescoping: sum of rewards = -kRelErr(last) since kPrev starts at 0 (burger_environment.py:128,175)
    from marlpde_b200 import _lib as LB
    kprev = env._get(LB.FIELD_KPREV, (B,), torch.float64)
    assert torch.allclose(tot[:, 0], -kprev, rtol=1e-9, atol=1e-12)
    # sampled rows of the full-size run (the HOT kernel variant the bench times) against the oracle, 1e-10
    from oracle.burger_oracle import BurgerOracle, forcing_tables, turbulence_ic
    from oracle.common import grid, spectral_rel_err
    rows = np.array([0, 1, 777, 1024, 2047, 2048, 3333, 4095])
    o = BurgerOracle(B=len(rows), L=TWO_PI, N=N, dt=1e-3, nu=0.02, forcing=True, dforce=False)
    o.setup_basis(32, "hat")
    tabs = [forcing_tables(int(s_), 5000) for s_ in seeds[rows]]
    o.set_forcing_tables(np.stack([t[0][:, :1] for t in tabs]), np.stack([t[1][:, :1] for t in tabs]))
    o.IC(u0=np.stack([turbulence_ic(grid(TWO_PI, N), TWO_PI, N, 0.0, int(s_)) for s_ in seeds[rows]]))
    prev = np.zeros(len(rows))
    for _ in range(3):
        for _ in range(10):
            o.step(acts[rows])
        err = np.array([spectral_rel_err(ref[o.ioutnum], o.Ek_ktt_row()[i], N) for i in range(len(rows))])
        o_rw, prev = prev - err, err
    u_gpu = env.u.cpu().numpy()[rows]
    assert np.max(np.abs(u_gpu - o.u)) <= 1e-10 * np.max(np.abs(o.u))
    assert np.max(np.abs(env.v.cpu().numpy()[rows] - o.v)) <= 1e-10 * np.max(np.abs(o.v))
    assert np.max(np.abs(st.cpu().numpy()[rows] - o.state())) <= 1e-9 * np.max(np.abs(o.state()))
    np.testing.assert_allclose(rw.cpu().numpy()[rows, 0], o_rw, rtol=1e-5, atol=1e-9)


def test_host_buffer_step_equals_device_step(golden):
    """mpde_step_host (pinned host in/out through HostPipeline) == the device-pointer path, bitwise."""
    from marlpde_b200.pipeline import HostPipeline
    g = golden("burger_steps.npz")
    V, A = g["eddy_forced/v"], g["eddy_forced/actions"]
    rows = [0, 5, 11, 17]
    ref_env, _ = make_env("eddy_forced", g, B=4, history=False)
    ref_env.IC(v0=V[rows])
    ref_env.set_spectrum_reference(np.abs(np.random.default_rng(0).normal(1, 0.1, (61, 16))) + 0.1)
    envs = []
    for _ in range(2):
        e, _m = make_env("eddy_forced", g, B=4, history=False)
        e.IC(v0=V[rows])
        e.set_spectrum_reference(np.abs(np.random.default_rng(0).normal(1, 0.1, (61, 16))) + 0.1)
        envs.append(e)
    pipe = HostPipeline(envs, 10)
    for k in range(2):
        pipe.submit(k, A[[0, 1, 2, 3]])
    st_ref, rw_ref = ref_env.step_n(A[[0, 1, 2, 3]], 10)
    for k in range(2):
        st, rw = pipe.collect(k)
        assert torch.equal(st, st_ref.cpu()) and torch.equal(rw, rw_ref.cpu())
    assert torch.equal(envs[1].v, ref_env.v)


@pytest.mark.parametrize("A,ver", [(1, 0), (4, 0), (32, 0), (4, 2), (8, 3), (32, 4)])
def test_multi_agent_mse_variant_equals_generic_kernel(A, ver):
    """history=False + MSE truth table selects the multi-agent training specialisation of the step kernel (no history /
    multi-column forcing / u_prev branches in the sub-step loop); it must reproduce the generic kernel's bits in state,
    per-agent reward and spectrum for every agent count and state layout."""
    from marlpde_b200 import Burger
    B, N = 7, 32
    rng = np.random.default_rng(100 * A + ver)
    truth = rng.normal(1.0, 0.3, (41, N))
    acts = rng.uniform(0.0, 0.02, (3, B, 32))
    out = []
    for history in (True, False):
        env = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=40, case="turbulence", forcing=False, dforce=False,
                     seed=42 + np.arange(B), version=ver, numAgents=A, nenvs=B, history=history)
        env.setup_basis(32, "hat")
        env.set_truth_table(truth[None])
        rows = []
        for s in range(3):
            st, rw = env.step_n(acts[s], 10)
            rows.append((st.clone(), rw.clone()))
        rows.append((env.v.clone(), env.getMseReward().clone()))        # nsub = 0: reward of the current state
        out.append(rows)
    for (a0, b0), (a1, b1) in zip(*out):
        assert torch.equal(a0, a1) and torch.equal(b0, b1)
