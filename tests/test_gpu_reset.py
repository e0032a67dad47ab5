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
