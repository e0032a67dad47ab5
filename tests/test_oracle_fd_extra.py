"""oracle/fd_extra_oracle.py (DiffusionError, Laplace) against golden vectors recorded from the reference classes.
These two solvers have no CUDA path yet: the pinned oracle is the first step of that row."""
import numpy as np
import pytest

from oracle.fd_extra_oracle import DiffusionErrorOracle, LaplaceOracle


@pytest.mark.parametrize("tag,agents", [("de_noact", 0), ("de_one", 1), ("de_marl", 32), ("de_box", 32)])
def test_diffusion_error_oracle(golden, tag, agents):
    g = golden("fd_extra.npz")
    U, A = g[f"{tag}/u"], g[f"{tag}/actions"]
    o = DiffusionErrorOracle(N=U.shape[1], dt=1e-3, nu=0.05)
    o.IC(U[0])
    for i in range(len(U) - 1):
        if agents == 0:
            o.step()
        else:
            o.step(A[i], agents)
        assert np.max(np.abs(o.u - U[i + 1])) <= 1e-13 * max(1.0, np.max(np.abs(U[i + 1]))), (tag, i)
    if tag != "de_box":
        sol = o.u0 * np.exp(-(2 * np.pi / o.L) ** 2 * o.nu * o.t)           # DiffusionError.py:288
        sec = o.N // max(agents, 1)
        rw = -((sol - o.u) ** 2).reshape(max(agents, 1), sec).mean(axis=1)
        np.testing.assert_allclose(rw, g[f"{tag}/reward"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(o.u, g[f"{tag}/state"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag", ["lp_sin", "lp_gauss"])
def test_laplace_oracle(golden, tag):
    g = golden("fd_extra.npz")
    U, A, R = g[f"{tag}/u"], g[f"{tag}/actions"], g[f"{tag}/reward"]
    o = LaplaceOracle(N=U.shape[1] - 1, dt=0.01)
    o.IC(U[0], g[f"{tag}/force"])
    nA = o.N - 1
    for i in range(len(U) - 1):
        o.step(A[i], nA)
        assert np.max(np.abs(o.u - U[i + 1])) <= 1e-12 * max(1.0, np.max(np.abs(U[i + 1]))), (tag, i)
        np.testing.assert_allclose(o.direct_reward(), R[i], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(o.state(nA), g[f"{tag}/state"], rtol=1e-12)
