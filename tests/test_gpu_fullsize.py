"""Full-size BASELINE configurations 3, 4, 5 through size-independent properties (rows of the big batch equal small
re-runs bitwise, formulas re-evaluated from the returned fields, oracle on a few rows), and parity of every team-size
variant of the Burgers kernel (16 / 8 / 4 lanes per environment use different FFT code paths)."""
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_burger import make_env, rel, TWO_PI

pytestmark = pytest.mark.gpu


@pytest.fixture
def team(monkeypatch):
    def set_team(ts):
        if ts is None:
            monkeypatch.delenv("MPDE_TS", raising=False)
        else:
            monkeypatch.setenv("MPDE_TS", str(ts))
    return set_team


@pytest.mark.parametrize("ts", [16, 8, 4, -8])
@pytest.mark.parametrize("case", ["eddy_forced", "direct", "dsm", "ssm_act"])
def test_every_team_size_matches_the_reference(golden, team, ts, case):
    """Free-running 60 steps against the reference's golden trajectory with the team size forced (MPDE_TS):
    shuffle FFT (16, 8 lanes) and 4 x 4 shared-memory-transposed FFT (4 lanes), 1e-10 on u, v, Fn_old."""
    team(ts)
    g = golden("burger_steps.npz")
    env, M = make_env(case, g)
    U, V, F, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/Fn_old"], g[f"{case}/actions"]
    env.IC(v0=V[0])
    worst = 0.0
    for i in range(len(U) - 1):
        env.step(A[i] if M else None)
        if i % 6 == 5 or i < 3:
            worst = max(worst, rel(env.v, V[i + 1]), rel(env.u, U[i + 1]), rel(env.Fn_old, F[i + 1]))
    assert worst < 1e-10, (case, ts, worst)


@pytest.mark.parametrize("ts", [16, 8, 4])
def test_every_team_size_is_batch_invariant(golden, team, ts):
    """Ragged batch (B = 37: partial warps and idle teams) with state + spectral reward: every row equals the same
    environment run alone, bitwise, for each team size."""
    team(ts)
    g = golden("burger_steps.npz")
    V, A = g["eddy_forced/v"], g["eddy_forced/actions"]
    B = 37
    rows = np.arange(B) % len(V)
    ref = np.abs(np.random.default_rng(0).normal(1, 0.1, (61, 16))) + 0.1
    env, _ = make_env("eddy_forced", g, B=B, history=False)
    env.IC(v0=V[rows]); env.set_spectrum_reference(ref)
    st, rw = env.step_n(A[rows % len(A)], 7)
    for j in (0, 1, 5, 17, 31, 32, 36):
        one, _ = make_env("eddy_forced", g, B=1, history=False)
        one.IC(v0=V[rows[j]]); one.set_spectrum_reference(ref)
        s1, r1 = one.step_n(A[rows[j] % len(A)], 7)
        assert torch.equal(s1[0], st[j]) and torch.equal(r1[0], rw[j]) and torch.equal(one.v, env.v[j]), (ts, j)


def test_full_size_config5_marl_mse(team):
    """BASELINE config 5 per GPU: 8192 envs x N=32, 32 per-gridpoint agents (state windows of 3, one action each),
    MSE reward against a shared truth table, nIntermediate = 10."""
    team(None)
    from marlpde_b200 import Burger
    B, N, A = 8192, 32, 32
    rng = np.random.default_rng(5)
    seeds = 42 + (np.arange(B) % 11)
    kw = dict(L=TWO_PI, N=N, dt=1e-3, nu=0.02, tend=1, case="turbulence", forcing=False, dforce=False, seed=seeds,
              version=0, numAgents=A, history=False)
    env = Burger(nenvs=B, **kw)
    env.setup_basis(32, "hat")
    truth = rng.normal(1.0, 0.3, (1001, N))
    env.set_truth_table(truth[None])
    acts = rng.uniform(0.0, 0.02, (B, 32))
    u0 = env.u.clone()
    st, rw = env.step_n(acts, 10)
    assert st.shape == (B, A * 3) and rw.shape == (B, A) and torch.isfinite(st).all() and torch.isfinite(rw).all()
    # state: agent a sees d2u/dx2 at points a-1, a, a+1 (Burger.py:657-670), re-evaluated from the returned field
    u = env.u
    dx = TWO_PI / N
    d2 = (torch.roll(u, 1, 1) - 2 * u + torch.roll(u, -1, 1)) / dx ** 2
    win = torch.stack([torch.roll(d2, 1, 1), d2, torch.roll(d2, -1, 1)], dim=2).reshape(B, A * 3)
    assert rel(st, win.cpu().numpy()) < 1e-9
    # rows of the big batch equal a small re-run, bitwise (batch invariance at full size)
    idx = [0, 1, 4095, 4096, 8191]
    small = Burger(nenvs=len(idx), **{**kw, "seed": seeds[idx]})
    small.setup_basis(32, "hat"); small.set_truth_table(truth[None])
    assert torch.equal(small.u, u0[idx])
    s2, r2 = small.step_n(acts[idx], 10)
    assert torch.equal(s2, st[idx]) and torch.equal(r2, rw[idx]) and torch.equal(small.v, env.v[idx])
    # one more single solver step: reward = -(truth[ioutnum] - u)^2 per point (A = N agents own one point each,
    # Burger.py:589-599)
    _, r1 = env.step_n(acts, 1)
    expect = -((torch.as_tensor(truth[11], device=env.device) - env.u) ** 2)
    assert rel(r1, expect.cpu().numpy()) < 1e-9


@pytest.mark.parametrize("case", ["eddy_forced", "direct", "ssm_act"])
def test_eight_and_four_lane_kernels_agree_bitwise(golden, team, case):
    """The radix-2^2 8-lane kernel (team_lanes=-8) and the 4-lane 4 x 4 shared-memory-transpose kernel perform the same
    arithmetic: identical bits in v, Fn_old, u, state and spectral reward, which is what lets team_lanes=-1 follow the
    batch size without changing any result."""
    team(None)
    g = golden("burger_steps.npz")
    V, A = g[f"{case}/v"], g[f"{case}/actions"]
    B = 13
    rows = np.arange(B) % len(V)
    ref = np.abs(np.random.default_rng(0).normal(1, 0.1, (61, 16))) + 0.1
    out = []
    for lanes in (-8, 4):       # -8: the radix-2^2 8-lane network
        env, M = make_env(case, g, B=B, history=False, team_lanes=lanes)
        env.IC(v0=V[rows]); env.set_spectrum_reference(ref)
        a = A[rows % len(A)] if M else None
        res = [env.step_n(a, 7), env.step_n(a, 3)]
        out.append((env.v.clone(), env.Fn_old.clone(), env.u.clone(), res[1][0].clone(), res[1][1].clone(), res[0][1].clone()))
    for x, y in zip(*out):
        assert torch.equal(x, y), case


def test_team_lanes_option_equals_env_override(golden, team):
    """Burger(team_lanes=4) selects the 4-lane kernels for that handle only: same bits as MPDE_TS=4, and the default
    handle next to it keeps the default team."""
    g = golden("burger_steps.npz")
    V, A = g["eddy_forced/v"], g["eddy_forced/actions"]
    team(None)
    a, _ = make_env("eddy_forced", g, B=5, history=False, team_lanes=4)
    d, _ = make_env("eddy_forced", g, B=5, history=False, team_lanes=-1)      # consistent auto: -8 at this batch size
    team(4)
    b, _ = make_env("eddy_forced", g, B=5, history=False)
    for e in (a, b, d):
        e.IC(v0=V[:5])
    b.step_n(A[:5], 9, want_reward=False)
    team(None)
    a.step_n(A[:5], 9, want_reward=False)
    d.step_n(A[:5], 9, want_reward=False)
    assert torch.equal(a.v, b.v)
    assert torch.equal(a.v, d.v)          # 4-lane and radix-2^2 8-lane kernels agree bitwise
    plain, _ = make_env("eddy_forced", g, B=5, history=False)                   # default: radix-2 8-lane network
    plain.IC(v0=V[:5]); plain.step_n(A[:5], 9, want_reward=False)
    assert rel(plain.v, a.v.cpu().numpy()) < 1e-12


def test_consistent_auto_team_is_batch_size_invariant(team):
    """team_lanes=-1: 4-lane kernels for the big batch, radix-2^2 8-lane kernels for the small re-run -- same bits."""
    team(None)
    from marlpde_b200 import Burger
    B, N = 6400, 32
    rng = np.random.default_rng(2)
    seeds = 42 + (np.arange(B) % 5)
    kw = dict(L=TWO_PI, N=N, dt=1e-3, nu=0.02, tend=1, case="turbulence", forcing=True, dforce=False, history=False, team_lanes=-1)
    ref = np.abs(rng.normal(1.0, 0.1, (1001, 16))) * 1e-3 + 1e-6
    acts = rng.uniform(0.0, 0.02, (B, 32))
    big = Burger(nenvs=B, seed=seeds, **kw)
    big.setup_basis(32, "hat"); big.set_spectrum_reference(ref)
    st, rw = big.step_n(acts, 10)
    idx = [0, 3, 3199, 6399]
    small = Burger(nenvs=len(idx), seed=seeds[idx], **kw)
    small.setup_basis(32, "hat"); small.set_spectrum_reference(ref)
    s2, r2 = small.step_n(acts[idx], 10)
    assert torch.equal(s2, st[idx]) and torch.equal(r2, rw[idx]) and torch.equal(small.v, big.v[idx])


def test_full_size_config3_ks(golden):
    """BASELINE config 3: 8192 KS environments, L = 22, N = 64, dt = 0.25, M = 64 hat basis."""
    from marlpde_b200 import KS
    from oracle.ks_oracle import KSOracle
    B, N, M = 8192, 64, 64
    rng = np.random.default_rng(3)
    u0 = rng.normal(0.0, 1e-3, (B, N))
    acts = rng.normal(0.0, 1e-3, (B, M))
    ks = KS(L=22, N=N, dt=0.25, nsteps=40, nenvs=B, u0=u0)
    ks.setup_basis(M, "hat")
    st, _ = ks.step_n(acts, 8, want_reward=False)
    assert st.shape == (B, 2 * N) and torch.isfinite(st).all()
    idx = [0, 3, 4097, 8191]
    small = KS(L=22, N=N, dt=0.25, nsteps=40, nenvs=len(idx), u0=u0[idx])
    small.setup_basis(M, "hat")
    s2, _ = small.step_n(acts[idx], 8, want_reward=False)
    assert torch.equal(s2, st[idx]) and torch.equal(small.v, ks.v[idx])
    o = KSOracle(B=len(idx), L=22, N=N, dt=0.25)
    o.setup_basis(M, "hat")
    o.IC(u0=u0[idx])
    for _ in range(8):
        o.step(acts[idx])
    assert rel(ks.v[idx], o.v) < 1e-10
    v = ks.v
    assert torch.equal(v[:, 1:N // 2], v[:, N // 2 + 1:].flip(1).conj())      # exact Hermitian symmetry


def test_full_size_config4_dns():
    """BASELINE config 4: 512 Burgers DNS environments at N = 1024 (CTA-resident kernel), u history every step."""
    from marlpde_b200 import Burger
    B, N, steps = 512, 1024, 40
    seeds = 100 + np.arange(B) % 5
    dns = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=steps, case="turbulence", seed=seeds, nenvs=B, history=True)
    assert dns.simulate() != -1
    uu = dns.uu
    assert uu.shape == (B, steps + 1, N) and torch.isfinite(uu).all()
    assert torch.equal(uu[0], uu[5]) and torch.equal(uu[3], uu[508])           # same seed -> same trajectory, bitwise
    one = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=steps, case="turbulence", seed=int(seeds[2]), history=True)
    one.simulate()
    assert torch.equal(one.uu.reshape(steps + 1, N), uu[2])
    # the history rows are Re ifft of the spectrum rows (complex64 history: 1e-6), energy decays without forcing
    vv = dns.vv[2].to(torch.complex128)
    assert rel(torch.fft.ifft(vv, dim=-1).real, uu[2].cpu().numpy()) < 1e-5
    e = (uu[2] - uu[2].mean(-1, keepdim=True)).pow(2).mean(-1)
    assert bool((e[1:] <= e[:-1] * (1 + 1e-12)).all())
