"""Device-side episode reset: DNS -> LES spectral hand-off and the 'turbulence' initial field generated
on the GPU, against the host (numpy, reference order of operations) formulas."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TWO_PI = 2 * np.pi


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _host_handoff(v0, k, offset, g):
    """burger_environment.py:110-111, literal."""
    v0off = v0 * np.exp(1j * 2 * np.pi * offset * k)
    return np.concatenate((v0off[:((g + 1) // 2)], v0off[-(g - 1) // 2:])) * g / len(v0)


@pytest.mark.parametrize("shifted", [False, True])
def test_handoff_matches_the_environment_formula(shifted):
    from marlpde_b200 import Burger
    from marlpde_b200.hostmath import fft_wavenumbers
    B, N, Nd = 37, 32, 512
    rng = np.random.default_rng(0)
    src = rng.normal(size=(3, Nd)) + 1j * rng.normal(size=(3, Nd))
    kd = fft_wavenumbers(TWO_PI, Nd)
    smap = np.arange(B) % 3
    off = rng.normal(0, 0.3, B) if shifted else np.zeros(B)
    a = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=10, case="zero", nenvs=B, history=False)
    b = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=10, case="zero", nenvs=B, history=False)
    a.IC_handoff(src, kd, src_map=smap, offsets=off if shifted else None)
    b.IC(v0=np.stack([_host_handoff(src[smap[e]], kd, off[e], N) for e in range(B)]))
    if shifted:
        assert rel(a.v, b.v.cpu().numpy()) < 1e-14 and rel(a.u, b.u.cpu().numpy()) < 1e-13
        assert rel(a.Fn_old, b.Fn_old.cpu().numpy()) < 1e-13
    else:       # no phase factor: exactly the same bits as the host path
        assert torch.equal(a.v, b.v) and torch.equal(a.u, b.u) and torch.equal(a.Fn_old, b.Fn_old)
    # masked reset: only the selected environments change
    a.step_n(None, 3, want_state=False, want_reward=False)
    before = a.v.clone()
    mask = (np.arange(B) % 2 == 0)
    a.IC_handoff(src, kd, src_map=smap, offsets=off if shifted else None, mask=mask)
    after = a.v
    assert torch.equal(after[1::2], before[1::2]) and rel(after[0::2], b.v[0::2].cpu().numpy()) < 1e-14


@pytest.mark.parametrize("N", [32, 512])
def test_turbulence_ic_on_device_matches_the_host_loop(N):
    from marlpde_b200 import Burger
    B = 24
    seeds = 7 + 3 * np.arange(B)
    off = np.random.default_rng(1).normal(0, 0.2, B)
    dev = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=5, case="zero", seed=seeds, nenvs=B, history=False, offset=off)
    host = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=5, case="zero", seed=seeds, nenvs=B, history=False, offset=off)
    dev.IC(case="turbulence", on_device=True)
    host.IC(case="turbulence", on_device=False)
    u = dev.u.cpu().numpy()
    assert rel(dev.u, host.u.cpu().numpy()) < 1e-12 and rel(dev.v, host.v.cpu().numpy()) < 1e-12
    rms = np.sqrt(np.mean((u - 1.0) ** 2, axis=1))
    assert np.all((rms > 0.6) & (rms < 0.8))                                   # Burger.py:259-260
    dev.step_n(None, 5, want_state=False, want_reward=False)
    host.step_n(None, 5, want_state=False, want_reward=False)
    assert rel(dev.v, host.v.cpu().numpy()) < 1e-11


def test_spline_table_on_device_matches_fitpack():
    """mpde_eval_spline_table against SciPy's FITPACK evaluation of the same spline, cubic and linear, for
    shifted + wrapped grids; then the MSE reward of a batch with 20 distinct offsets (device-sampled truth tables) against
    the same batch fed host-sampled tables."""
    from marlpde_b200 import Burger
    from marlpde_b200.hostmath import TruthInterpolant
    rng = np.random.default_rng(4)
    L, Nd, rows, N = TWO_PI, 128, 41, 32
    xd = np.linspace(0, L, Nd, endpoint=False)
    tt = np.concatenate(([0.], np.cumsum(np.full(rows - 1, 1e-3))))
    uu = 1.0 + np.sin(2 * xd[None, :] + 40 * tt[:, None]) + 0.05 * rng.normal(size=(rows, Nd))
    x = np.linspace(0, L, N, endpoint=False)
    shifts = rng.normal(0, 0.4, 20)
    grids = []
    for sh in shifts:
        g = x + sh
        g[g > L] -= L
        g[g < 0] += L
        grids.append(g)
    grids = np.stack(grids)
    for kind in ("cubic", "linear"):
        f = TruthInterpolant(xd, tt, uu, kind=kind)
        dev = f.rows_device(grids, tt, torch.device("cuda", 0), torch.float64).cpu().numpy()
        host = np.stack([f.rows(g, tt) for g in grids])
        assert np.max(np.abs(dev - host)) < 1e-12 * np.max(np.abs(host)), kind
    B = 20
    kw = dict(L=L, N=N, dt=1e-3, nu=0.02, nsteps=rows - 1, case="sinus", seed=3, nenvs=B, history=False, offset=shifts, numAgents=4)
    a, b = Burger(**kw), Burger(**kw)
    a.setGroundTruth(xd, tt, uu)
    a._ensure_truth(shifts)                                   # > 8 distinct shifts: sampled on the device
    f = TruthInterpolant(xd, tt, uu, kind="cubic")
    b.set_truth_table(np.stack([f.rows(g, tt) for g in grids]), env_map=np.arange(B, dtype=np.int32))
    for env in (a, b):
        env.setup_basis(8, "hat")
    acts = rng.uniform(-0.1, 0.1, (B, 8))
    _, ra = a.step_n(acts, 5)
    _, rb = b.step_n(acts, 5)
    assert rel(ra, rb.cpu().numpy()) < 1e-10


@pytest.mark.parametrize("nsteps,stepper,nunoise,nseeds", [(1000, 1, False, 4096), (5000, 1, False, 48), (777, 4, True, 96), (300, 20, True, 33)])
def test_device_forcing_tables_match_numpy_stream(nsteps, stepper, nunoise, nseeds):
    """SURVEY 8f-1 / Burger.py:66,88-95: the device generator (one warp per seed: MT19937 init_genrand, 53-bit doubles,
    legacy polar gauss) reproduces rows 1..3, columns < stepper of `np.random.seed(seed); [uniform();] normal(size=(32,nsteps)) x 2`.
    Acceptance of the polar candidates is exact arithmetic (identical stream position); the kept value goes through log(),
    whose last bit differs between CUDA and glibc: <= 2 ulp allowed, and most entries must be bit-equal."""
    from marlpde_b200.Burger import device_forcing_tables
    seeds = 42 + np.arange(nseeds)
    nu, r1, r2 = device_forcing_tables(seeds, nsteps, stepper, nunoise, "cuda:0")
    e1, e2, en = np.empty_like(r1), np.empty_like(r2), np.empty(nseeds)
    for i, sd in enumerate(seeds):
        rs = np.random.RandomState(int(sd))
        if nunoise:
            en[i] = 0.01 + 0.02 * rs.uniform()
        e1[i] = rs.normal(loc=0., scale=1., size=(32, nsteps))[1:4, :stepper]
        e2[i] = rs.normal(loc=0., scale=1., size=(32, nsteps))[1:4, :stepper]
    if nunoise:
        assert np.array_equal(nu, en)
    for got, exp in ((r1, e1), (r2, e2)):
        assert np.all(np.abs(got - exp) <= 2 * np.spacing(np.abs(exp)))
        assert np.mean(got == exp) > 0.8


def test_burger_batch_with_per_env_seeds_uses_device_tables_and_matches_single_envs():
    """bench workload: seed = 42 + e per environment.  The batch built from device-generated tables steps
    like single environments built from the host NumPy stream (1e-12: the table entries agree to <= 2 ulp)."""
    from marlpde_b200 import Burger
    B, N, M = 64, 32, 32
    seeds = 42 + np.arange(B)
    kw = dict(N=N, dt=1e-3, nu=0.02, tend=0.2, case="turbulence", forcing=True, dforce=False, history=False)
    env = Burger(seed=seeds, nenvs=B, **kw)
    assert env._tables_on_device
    env.setup_basis(M, "hat")
    a = np.random.default_rng(0).uniform(0.0, 0.05, (B, M))
    env.step_n(a, 20)
    u = env.u.cpu().numpy()
    for e in (0, 17, 63):
        one = Burger(seed=int(seeds[e]), nenvs=1, **kw)
        assert not one._tables_on_device
        one.setup_basis(M, "hat")
        one.step_n(a[e:e + 1], 20)
        ref = one.u.cpu().numpy()
        assert np.max(np.abs(u[e] - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_masked_reset_keeps_per_env_counters_and_flags_the_host_scalars():
    """IC(mask=...) resets a subset.  The device counters are per environment (the reset ones restart at 0, the
    others keep running); the host scalars would describe neither, so history / checkpoint views refuse until a full reset."""
    from marlpde_b200 import Burger
    B, N = 6, 32
    env = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=40, case="turbulence", seed=42, nenvs=B, history=True)
    env.step_n(None, 7, want_state=False, want_reward=False)
    mask = np.array([1, 0, 0, 1, 0, 0], dtype=np.uint8)
    env.IC(case="turbulence", mask=mask)
    assert env.ioutnum_all.cpu().tolist() == [0, 7, 7, 0, 7, 7]
    with pytest.raises(RuntimeError, match="masked reset"):
        env.compute_Ek()
    with pytest.raises(RuntimeError, match="masked reset"):
        env.state_dict()
    env.step_n(None, 3, want_state=False, want_reward=False)
    assert env.ioutnum_all.cpu().tolist() == [3, 10, 10, 3, 10, 10]
    fresh = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=40, case="turbulence", seed=42, nenvs=1, history=False)
    fresh.step_n(None, 3, want_state=False, want_reward=False)
    assert torch.equal(env.v[0], fresh.v) and torch.equal(env.v[3], fresh.v)      # the reset rows restarted exactly
    env.IC(case="turbulence")                                                       # a full reset clears the flag
    env.compute_Ek()


@pytest.mark.parametrize("kind,mx,mt", [("cubic", 512, 201), ("cubic", 64, 33), ("linear", 32, 17), ("cubic", 9, 8)])
def test_spline_fit_on_device_matches_fitpack(kind, mx, mt):
    """mpde_fit_spline: knots identical to FITPACK's, coefficients of the interpolating spline equal to
    RectBivariateSpline(s = 0) -- what interp2d builds in setGroundTruth (Burger.py:322-323) -- to rounding."""
    from scipy.interpolate import RectBivariateSpline
    from marlpde_b200.hostmath import TruthInterpolant
    rng = np.random.default_rng(8)
    x = np.linspace(0, TWO_PI, mx, endpoint=False)
    t = np.concatenate(([0.], np.cumsum(np.full(mt - 1, 1e-3))))
    uu = np.sin(x[None, :] + 30 * t[:, None]) + 0.1 * rng.normal(size=(mt, mx))
    f = TruthInterpolant(x, t, uu, kind=kind)
    _, tx, ty, c = f.fit_device(torch.device("cuda", 0))
    k = 3 if kind == "cubic" else 1
    TX, TY, C = RectBivariateSpline(x, t, uu.T, kx=k, ky=k, s=0).tck
    assert np.array_equal(tx.cpu().numpy(), TX) and np.array_equal(ty.cpu().numpy(), TY)
    assert np.max(np.abs(c.cpu().numpy() - C)) <= 1e-12 * np.max(np.abs(C))
