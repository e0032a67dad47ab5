"""GPU parity of the DNS-size solvers (CTA-per-environment kernels, N = 256..2048, plus the
warp kernels at N = 128/256) against golden vectors recorded from the reference Burger class:
simulate() rows, final spectra, and the float32 Ek_ktt history (compute_Ek)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["turb1024", "sinus512", "turb256_forced", "forced_L100", "turb128", "turb2048"]


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("tag", CASES)
def test_burgers_dns_simulate(golden, tag):
    from marlpde_b200 import Burger
    g = golden("burger_dns.npz")
    N, L, dt, nsteps, forcing, st = g[tag + "/cfg"]
    N, nsteps, st = int(N), int(nsteps), int(st)
    dns = Burger(L=L, N=N, dt=dt, nu=0.02, nsteps=nsteps, u0=g[tag + "/u0"], forcing=bool(forcing), s=st, history=True)
    if forcing:
        dns.randfac1, dns.randfac2 = g[tag + "/randfac1"], g[tag + "/randfac2"]
    assert dns.simulate() is None
    assert dns.ioutnum == nsteps
    uu = dns.uu.cpu().numpy()
    assert rel(uu[::20], g[tag + "/rows"]) < 1e-10
    assert rel(dns.v, g[tag + "/v_final"]) < 1e-10
    assert rel(dns.Fn_old, g[tag + "/Fn_old_final"]) < 1e-10
    dns.compute_Ek()
    ektt = dns.Ek_ktt.cpu().numpy()
    # modes that are pure round-off of the FFT (sinus IC: ~1e-33) are not comparable between FFT libraries
    np.testing.assert_allclose(ektt[::20, :64], g[tag + "/Ek_ktt"], rtol=1e-5, atol=1e-20 * g[tag + "/Ek_ktt"].max())
    np.testing.assert_allclose(dns.Ek_tt.cpu().numpy()[::20], g[tag + "/Ek_tt"], rtol=1e-5)
    # tt accumulates t += dt exactly like the reference
    assert dns.tt[-1] == pytest.approx(nsteps * dt, rel=1e-12)


def test_dns_batch_of_seeds_matches_single_runs(golden):
    """B independent DNS in one launch == the same DNS alone (bitwise)."""
    from marlpde_b200 import Burger
    seeds = [42, 43, 44]
    batch = Burger(N=512, dt=1e-3, nu=0.02, nsteps=50, case="turbulence", seed=seeds, nenvs=3, history=False)
    batch.simulate()
    for i, s in enumerate(seeds):
        one = Burger(N=512, dt=1e-3, nu=0.02, nsteps=50, case="turbulence", seed=s, history=False)
        one.simulate()
        assert torch.equal(one.v, batch.v[i])


def test_dns_feeds_les_spectral_reward(golden):
    """End-to-end ground-truth path (burger_environment.py:11-16, 99-112, 172-176): a GPU DNS provides
    the forcing tables, the truncated IC and the Ek_ktt reference of a batch of LES environments."""
    from marlpde_b200 import Burger
    from oracle.burger_oracle import BurgerOracle, truncated_ic
    from oracle.common import spectral_rel_err
    T, dt, g_ = 0.2, 1e-3, 32
    dns = Burger(N=512, dt=dt, nu=0.02, tend=T, case="turbulence", forcing=True, seed=50, history=True)
    dns.simulate()
    dns.compute_Ek()
    B = 6
    les = Burger(N=g_, dt=dt, nu=0.02, tend=T, case="zero", forcing=True, dforce=False, seed=50, nenvs=B, history=False)
    les.randfac1, les.randfac2 = dns.randfac1, dns.randfac2
    les.setup_basis(32, "hat")
    v0 = truncated_ic(dns.v0.cpu().numpy(), dns.k, 0.0, g_)
    les.IC(v0=v0)
    les.set_spectrum_reference(dns)
    rng = np.random.default_rng(3)
    acts = rng.uniform(0.0, 0.03, (B, 32))
    o = BurgerOracle(B=B, N=g_, dt=dt, nu=0.02, forcing=True, dforce=False)
    o.setup_basis(32, "hat")
    o.set_forcing_tables(dns.randfac1, dns.randfac2)
    o.IC(v0=np.broadcast_to(v0, (B, g_)))
    ref = dns.Ek_ktt.cpu().numpy()
    prev = np.zeros(B)
    for s in range(10):
        st, rw = les.step_n(acts, 10)
        for _ in range(10):
            o.step(acts)
        err = spectral_rel_err(ref[o.ioutnum], o.Ek_ktt_row(), g_)
        np.testing.assert_allclose(rw[:, 0].cpu().numpy(), prev - err, rtol=1e-5, atol=1e-9)
        prev = err
        assert rel(st, o.state()) < 1e-10
