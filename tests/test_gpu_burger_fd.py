"""GPU parity of the batched finite-difference Burgers solver (marlpde_b200.Burger_fd) against golden vectors recorded
from the reference's Burger_fd class: u and v = fft(u) at every recorded step (1e-10), state, running spectrum, fused
sub-steps and batch invariance."""
import numpy as np
import pytest
import torch

from tests.test_oracle_burger_fd import FD_CASES

pytestmark = pytest.mark.gpu
TWO_PI = 2 * np.pi


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def make(case, g, B=1, **extra):
    from marlpde_b200 import Burger_fd
    kw = dict(FD_CASES[case])
    N, M = kw.pop("N", 32), kw.pop("M", 32)
    stepper = kw.pop("stepper", 1)
    env = Burger_fd(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=60, case="zero", s=stepper, nenvs=B, **kw, **extra)
    if M:
        env.setup_basis(M, "hat")
    if kw.get("forcing"):
        env.randfac1, env.randfac2 = g[f"{case}/randfac1"], g[f"{case}/randfac2"]
    return env, M


@pytest.mark.parametrize("case", sorted(FD_CASES))
def test_free_running_vs_reference(golden, case):
    g = golden("burger_fd.npz")
    U, V, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/actions"]
    env, M = make(case, g)
    env.IC(u0=U[0])
    worst = 0.0
    for i in range(len(U) - 1):
        env.step(A[i] if M else None)
        if i % 6 == 5 or i < 3:
            worst = max(worst, rel(env.u, U[i + 1]), rel(env.v, V[i + 1]))
    assert worst < 1e-10, (case, worst)
    assert rel(env.getState(as_tensor=True), g[f"{case}/state"]) < 1e-9
    N = env.N
    np.testing.assert_allclose(env.Ek_ktt_row().cpu().numpy().reshape(-1)[1:N // 2], g[f"{case}/Ek_ktt"][1:N // 2], rtol=2e-5)


def test_fused_substeps_and_batch_invariance(golden):
    g = golden("burger_fd.npz")
    case = "fd_eddy_forced"
    U, A = g[f"{case}/u"], g[f"{case}/actions"]
    rows = [0, 7, 19, 33, 41]
    big, _ = make(case, g, B=len(rows), history=False)
    big.IC(u0=U[rows])
    acts = A[[0, 10, 20, 30, 40]]
    st, _ = big.step_n(acts, 10, want_reward=False)
    for j, r in enumerate(rows):
        one, _ = make(case, g, B=1, history=False)
        one.IC(u0=U[r])
        for _ in range(10):
            one.step(acts[j])
        assert torch.equal(one.u, big.u[j]) and torch.equal(one.v, big.v[j])
        assert torch.equal(one.getState(as_tensor=True).reshape(-1), st[j])
