"""oracle/burger_fd_oracle.py against the golden vectors recorded from the reference's Burger_fd class."""
import numpy as np
import pytest

from oracle.burger_fd_oracle import BurgerFdOracle

FD_CASES = {
    "fd_noact": dict(M=0),
    "fd_direct": dict(dforce=True),
    "fd_eddy": dict(dforce=False),
    "fd_eddy_forced": dict(dforce=False, forcing=True),
    "fd_ssm": dict(ssm=True, M=0),
    "fd_ssmforce": dict(dforce=True, ssmforce=True),
    "fd_sinus64": dict(N=64, dforce=False, M=16),
    "fd_forced_s4": dict(forcing=True, stepper=4, dforce=True),
    "fd_dsm": dict(dsm=True, M=0),
    "fd_dsm_eddy": dict(dsm=True, dforce=False),
    "fd_v1": dict(dforce=False, version=1),
    "fd_v2": dict(dforce=True, version=2),
}


def make_oracle(case, g):
    kw = dict(FD_CASES[case])
    N, M = kw.pop("N", 32), kw.pop("M", 32)
    o = BurgerFdOracle(B=1, L=2 * np.pi, N=N, dt=1e-3, nu=0.02, **kw)
    if M:
        o.setup_basis(M, "hat")
    if kw.get("forcing"):
        o.set_forcing_tables(g[f"{case}/randfac1"], g[f"{case}/randfac2"])
    return o, M


@pytest.mark.parametrize("case", sorted(FD_CASES))
def test_fd_oracle_matches_reference(golden, case):
    g = golden("burger_fd.npz")
    U, V, A = g[f"{case}/u"], g[f"{case}/v"], g[f"{case}/actions"]
    o, M = make_oracle(case, g)
    o.IC(u0=U[0][None])
    worst = 0.0
    for i in range(len(U) - 1):
        o.step(A[i][None] if M else None)
        worst = max(worst, np.max(np.abs(o.u[0] - U[i + 1])) / np.max(np.abs(U[i + 1])),
                    np.max(np.abs(o.v[0] - V[i + 1])) / np.max(np.abs(V[i + 1])))
    assert worst < 1e-12, (case, worst)
    assert np.max(np.abs(o.state()[0] - g[f"{case}/state"])) <= 1e-9 * np.max(np.abs(g[f"{case}/state"]))
    N = o.N
    ek = o.Ek_ktt_row()[0]
    ref = g[f"{case}/Ek_ktt"]
    np.testing.assert_allclose(ek[1:N // 2], ref[1:N // 2], rtol=2e-5)
