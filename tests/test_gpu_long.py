"""north_star: "energy spectra after 1000 steps agreeing within 1 %" -- the CUDA path (fp64 AND fp32) free-running for
1000 solver steps against goldens recorded from the REAL reference (tests/golden/make_golden_long.py), and KS.getReward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROWS = (250, 500, 750, 1000)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag", ["eddy_forced", "eddy", "direct_forced"])
def test_burgers_spectrum_after_1000_steps(golden, tag, dtype):
    from marlpde_b200 import Burger
    g = golden("long_runs.npz")
    p = f"burger_{tag}/"
    seed, forcing, dforce, nsteps, hold, N, M = g[p + "cfg"]
    nsteps, hold, N, M = int(nsteps), int(hold), int(N), int(M)
    env = Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=nsteps, case="zero", forcing=bool(forcing), dforce=bool(dforce),
                 nenvs=3, dtype=dtype, history=False)          # 3 copies: rows of a batch, must be identical
    env.setup_basis(M, "hat")
    if forcing:
        env.randfac1, env.randfac2 = g[p + "randfac1"], g[p + "randfac2"]
    env.IC(u0=g[p + "u0"])
    A = g[p + "actions"]
    tol = 1e-5 if dtype == torch.float64 else 1e-2              # f64: float32 spectrum chain (Q6); f32: north_star 1 %
    for r in range(nsteps // hold):
        env.step_n(np.tile(A[r], (3, 1)), hold, want_state=False, want_reward=False)
        if (r + 1) * hold in ROWS:
            ref = g[p + "Ek_ktt"][ROWS.index((r + 1) * hold)]
            got = env.Ek_ktt_row().cpu().numpy()[:, :N // 2]
            assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[2])
            assert np.max(np.abs(got[0] - ref) / ref) < tol, (r + 1) * hold
    assert int((env.status != 0).sum()) == 0
    u = env.u.double().cpu().numpy()[0]
    uref = g[p + "u_final"]
    assert np.max(np.abs(u - uref)) <= (1e-9 if dtype == torch.float64 else 1e-2) * np.max(np.abs(uref))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_ks_spectrum_after_1000_steps(golden, dtype):
    from marlpde_b200 import KS
    g = golden("long_runs.npz")
    N, M, nsteps, hold = (int(x) for x in g["ks_n64/cfg"])
    ks = KS(L=22.0, N=N, dt=0.25, nsteps=nsteps, v0=g["ks_n64/v0"], nenvs=1, dtype=dtype, history=False)
    ks.setup_basis(M, "hat")
    A = g["ks_n64/actions"]
    for r in range(nsteps // hold):
        ks.step_n(A[r][None], hold, want_state=False, want_reward=False)
        if (r + 1) * hold in ROWS:
            ref = g["ks_n64/Ek_ktt"][ROWS.index((r + 1) * hold)]
            got = ks.Ek_ktt_row().cpu().numpy().reshape(-1)[:N // 2]
            err = np.max(np.abs(got[1:] - ref[1:]) / ref[1:])
            if dtype == torch.float64:
                assert err < 1e-2, ((r + 1) * hold, err)
            elif (r + 1) * hold <= 500:
                # fp32 on a chaotic attractor: trajectories decorrelate after ~150 time units (Lyapunov time ~ 10), so only
                # the rows before that are a numerics statement; the later ones are statistics of two different orbits
                assert err < 1e-2, ((r + 1) * hold, err)
            else:
                assert err < 0.25, ((r + 1) * hold, err)


def test_ks_get_reward_matches_reference(golden):
    """KS.getReward (KS.py:360-367): -|u - f_truth(x, t)| on the float32 real-space row, truth = cubic spline of a DNS."""
    from marlpde_b200 import KS
    g = golden("long_runs.npz")
    ks = KS(L=22.0, N=64, dt=0.25, nsteps=40, v0=g["ks_reward/v0"], nenvs=1)
    ks.setup_basis(16, "hat")
    ks.setGroundTruth(g["ks_reward/dns_tt"], g["ks_reward/dns_x"], g["ks_reward/dns_uu"])
    A = g["ks_reward/actions"]
    for r in range(3):
        ks.step_n(A[r][None], 4, want_state=False, want_reward=False)
        got = np.asarray(ks.getReward(), dtype=np.float64).reshape(-1)
        np.testing.assert_allclose(got, g["ks_reward/rewards"][r], rtol=2e-5, atol=2e-6)
