"""GPU parity of the Diffusion / Advection stencil environments against golden vectors recorded
from the reference classes (actions as stencil weights, per-agent windows, MSE / direct rewards,
analytic solutions, implicit Euler)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


DIFF = {"plain": ("sinus", 1), "lap": ("sinus", 1), "point_A1": ("box", 1), "point_A4": ("gaussian", 4),
        "point_AN": ("sinus", 32), "implicit": ("box", 1)}


@pytest.mark.parametrize("tag", sorted(DIFF))
def test_diffusion(golden, tag):
    from marlpde_b200 import Diffusion
    g = golden("fd.npz")
    p = f"diff_{tag}/"
    N, L, dt, nu, A = g[p + "cfg"]
    N, A = int(N), int(A)
    case = DIFF[tag][0]
    U, acts, S = g[p + "u"], g[p + "actions"], g[p + "states"]
    d = Diffusion(L=L, N=N, dt=dt, nu=nu, nsteps=30, case=case, implicit=(tag == "implicit"))
    assert rel(d.u, U[0]) < 1e-15
    assert rel(np.array(d.getState(A)).reshape(-1), S[0].reshape(-1)) < 1e-15
    for i in range(len(U) - 1):
        if len(acts):
            a = acts[i]
            d.step(a.tolist() if A == 1 else a.reshape(A, -1).tolist(), numAgents=A)
        else:
            d.step()
        assert rel(d.u, U[i + 1]) < 1e-12, (tag, i)
        assert rel(np.array(d.getState(A)).reshape(-1), S[i + 1].reshape(-1)) < 1e-12
        if len(g[p + "mse"]):
            np.testing.assert_allclose(np.atleast_1d(d.getMseReward(A)), g[p + "mse"][i], rtol=1e-8, atol=1e-30)
        if len(g[p + "direct"]):
            np.testing.assert_allclose(d.getDirectReward(A), g[p + "direct"][i], rtol=1e-8, atol=1e-15)
    assert rel(d.uu, U) < 1e-12
    if case == "sinus":
        assert rel(d.solution, g[p + "solution"]) < 1e-14


@pytest.mark.parametrize("tag,A", [("lax", 1), ("global", 1), ("point_A1", 1), ("point_A4", 4)])
def test_advection(golden, tag, A):
    from marlpde_b200 import Advection
    g = golden("fd.npz")
    p = f"adv_{tag}/"
    N, L, dt, nu, _ = g[p + "cfg"]
    N = int(N)
    U, acts, S = g[p + "u"], g[p + "actions"], g[p + "states"]
    a_ = Advection(L=L, N=N, dt=dt, nu=nu, nsteps=30, case="sinus")
    for i in range(len(U) - 1):
        if len(acts):
            a = acts[i]
            a_.step(a.tolist() if A == 1 else a.reshape(A, -1).tolist(), numAgents=A)
        else:
            a_.step()
        assert rel(a_.u, U[i + 1]) < 1e-12, (tag, i)
        assert rel(np.array(a_.getState(A)).reshape(-1), S[i + 1].reshape(-1)) < 1e-12
        np.testing.assert_allclose(np.atleast_1d(a_.getMseReward(A)), g[p + "mse"][i], rtol=1e-8, atol=1e-30)
    assert rel(a_.solution, g[p + "solution"]) < 1e-14


def test_known_answers_and_batching():
    """Action -2 == standard Laplacian; Lax weights == FDstep; batch rows == single envs (bitwise);
    convergence to the analytic sinus solutions."""
    from marlpde_b200 import Diffusion, Advection
    N, L = 64, 2 * np.pi
    d0 = Diffusion(L=L, N=N, dt=1e-3, nu=0.1, nsteps=200, case="sinus")
    d1 = Diffusion(L=L, N=N, dt=1e-3, nu=0.1, nsteps=200, case="sinus")
    d0.simulate()
    for _ in range(200):
        d1.step([-2.0])
    assert rel(d0.u, d1.u.cpu().numpy()) < 1e-13
    assert rel(d0.u, d0.getAnalyticalSolution(d0.t)) < 1e-3
    offs = np.array([0.0, 0.3, -0.2, 0.1, 0.7])
    db = Diffusion(L=L, N=N, dt=1e-3, nu=0.1, nsteps=50, case="sinus", nenvs=5, offset=offs)
    acts = -2.0 + 0.1 * np.random.default_rng(0).normal(size=(5, N))
    db.step_n(acts, 1, 50)
    for e in (0, 2, 4):
        d = Diffusion(L=L, N=N, dt=1e-3, nu=0.1, nsteps=50, case="sinus", offset=offs[e])
        d.step_n(acts[e], 1, 50)
        assert torch.equal(d.u, db.u[e])
    a0 = Advection(L=L, N=N, dt=0.09, nu=1.0, nsteps=20, case="sinus")
    a1 = Advection(L=L, N=N, dt=0.09, nu=1.0, nsteps=20, case="sinus")
    al = a0.alpha
    a0.simulate()
    for _ in range(20):
        a1.step([0.5 + 0.5 * al, 0.5 - 0.5 * al])
    assert rel(a0.u, a1.u.cpu().numpy()) < 1e-13
    assert rel(a0.u, a0.getAnalyticalSolution(a0.t)) < 2e-2


@pytest.mark.parametrize("tag,agents", [("de_noact", 0), ("de_one", 1), ("de_marl", 32), ("de_box", 32)])
def test_diffusion_error(golden, tag, agents):
    """DiffusionError.step / getMseReward / getState (DiffusionError.py:160-298) against goldens recorded from the reference
    class; batch of 3 identical environments must agree bitwise."""
    from marlpde_b200 import DiffusionError
    g = golden("fd_extra.npz")
    U, A = g[f"{tag}/u"], g[f"{tag}/actions"]
    case = "box" if tag == "de_box" else "sinus"
    d = DiffusionError(L=2 * np.pi, N=U.shape[1], dt=1e-3, nu=0.05, nsteps=len(U) - 1, case=case, nenvs=3)
    assert np.array_equal(d.u.cpu().numpy()[0], U[0])
    for i in range(len(U) - 1):
        if agents == 0:
            d.step()
        else:
            d.step(np.tile(A[i].reshape(1, -1), (3, 1)), agents)
        u = d.u.cpu().numpy()
        assert np.array_equal(u[0], u[1]) and np.array_equal(u[0], u[2])
        assert np.max(np.abs(u[0] - U[i + 1])) <= 1e-12 * max(1.0, np.max(np.abs(U[i + 1]))), (tag, i)
    if case == "sinus":
        rw = d.getMseReward(max(agents, 1)).cpu().numpy()[0]
        np.testing.assert_allclose(rw, g[f"{tag}/reward"], rtol=1e-8, atol=1e-18)
    np.testing.assert_allclose(d.getState(1).cpu().numpy()[0], g[f"{tag}/state"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag,sforce,ic", [("lp_sin", "sin", "one"), ("lp_gauss", "gaussian", "sin")])
def test_laplace(golden, tag, sforce, ic):
    """Laplace.step / getDirectReward / getState (Laplace.py:116-166) against goldens recorded from the reference class."""
    from marlpde_b200 import Laplace
    g = golden("fd_extra.npz")
    U, A, R = g[f"{tag}/u"], g[f"{tag}/actions"], g[f"{tag}/reward"]
    lp = Laplace(L=2 * np.pi, N=U.shape[1] - 1, dt=0.01, ic=ic, sforce=sforce, episodeLength=len(U) - 1, nenvs=2)
    nA = lp.N - 1
    assert np.max(np.abs(lp.u.cpu().numpy()[0] - U[0])) <= 1e-15 and np.max(np.abs(lp.force[0] - g[f"{tag}/force"])) <= 1e-15
    for i in range(len(U) - 1):
        _, rw = lp.step_n(np.tile(A[i].reshape(1, -1), (2, 1)), nA, 1, want_reward=True)
        u = lp.u.cpu().numpy()
        assert np.array_equal(u[0], u[1])
        assert np.max(np.abs(u[0] - U[i + 1])) <= 1e-12 * max(1.0, np.max(np.abs(U[i + 1]))), (tag, i)
        np.testing.assert_allclose(rw.cpu().numpy()[0], R[i], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(np.asarray(lp.getDirectReward(nA).cpu().numpy()[0]), R[i], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(lp.getState(nA).cpu().numpy()[0], g[f"{tag}/state"], rtol=1e-12, atol=1e-13)
    assert np.array_equal(lp.uu.cpu().numpy()[0, -1], u[0])
