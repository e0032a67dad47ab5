"""CPU-only checks of the host-side logic and of the C-ABI library (load + exported symbols;
no compute calls -- those need a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from marlpde_b200 import _lib as LB
from marlpde_b200 import hostmath as hm
from oracle.burger_oracle import BurgerOracle
from oracle.common import fft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "marlpde_b200.h")).read()
    declared = set(re.findall(r"\b(mpde_[a-z0-9_]+)\s*\(", header))
    declared -= {"mpde_config", "mpde_env"}
    assert declared == set(LB.SIGNATURES), declared ^ set(LB.SIGNATURES)
    lib = LB.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mpde_abi_version() == LB.ABI_VERSION


def test_config_struct_matches_header_layout():
    assert ctypes.sizeof(LB.MpdeConfig) == 4 * 4 + 8 + 8 * 4 + 16


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = LB.lib()
    cfg = LB.MpdeConfig()
    cfg.struct_size = ctypes.sizeof(LB.MpdeConfig)
    cfg.nenvs, cfg.N, cfg.num_agents, cfg.stepper, cfg.L, cfg.dt = 4, 32, 1, 1, 6.28, 1e-3
    h = ctypes.c_void_p()
    assert lib.mpde_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.mpde_last_error()
    from marlpde_b200 import Burger
    with pytest.raises(RuntimeError, match="no CPU path"):
        Burger(N=32, case="zero")


def test_argument_validation_messages():
    lib = LB.lib()
    cfg = LB.MpdeConfig()
    h = ctypes.c_void_p()
    assert lib.mpde_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"size mismatch" in lib.mpde_last_error()
    cfg.struct_size = ctypes.sizeof(LB.MpdeConfig)
    cfg.nenvs, cfg.N, cfg.num_agents, cfg.stepper, cfg.L, cfg.dt = 4, 48, 1, 1, 6.28, 1e-3
    assert lib.mpde_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"power-of-two" in lib.mpde_last_error()
    cfg.N, cfg.num_agents = 32, 5
    assert lib.mpde_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"num_agents" in lib.mpde_last_error()


def test_basis_and_ic_match_golden(golden):
    g = golden("burger_steps.npz")
    x = hm.grid_points(2 * np.pi, 32)
    assert np.array_equal(x, g["direct/x"])
    assert np.array_equal(hm.fft_wavenumbers(2 * np.pi, 32), g["direct/k"])
    assert np.array_equal(hm.make_basis(x, 2 * np.pi, 32, "hat"), g["direct/basis"])
    assert np.array_equal(hm.make_basis(x, 2 * np.pi, 8, "uniform"), g["uniform8/basis"])
    assert np.array_equal(hm.make_basis(x, 2 * np.pi, 5, "hat"), g["hat5/basis"])
    assert np.array_equal(hm.make_basis(x, 2 * np.pi, 1, "hat"), g["one_action/basis"])
    assert np.array_equal(hm.turbulence_field(x, 2 * np.pi, 32, 0.0, 42), g["direct/u"][0])
    d = golden("burger_dns.npz")
    x1k = hm.grid_points(2 * np.pi, 1024)
    assert np.array_equal(hm.turbulence_field(x1k, 2 * np.pi, 1024, 0.0, 42), d["turb1024/u0"])


def test_forced_ic_uses_stream_after_tables(golden):
    d = golden("burger_dns.npz")
    rs = np.random.RandomState(42)
    rs.normal(size=(32, 120)); rs.normal(size=(32, 120))
    u0 = hm.forced_field(hm.grid_points(100.0, 256), 100.0, 256, rs)
    assert np.array_equal(u0, d["forced_L100/u0"])


@pytest.mark.parametrize("stepper,per_env", [(1, False), (4, False), (3, True)])
def test_forcing_coefficients_are_the_fft_of_the_forcing(stepper, per_env):
    B, N, L, dt = 3, 32, 2 * np.pi, 1e-3
    rng = np.random.default_rng(1)
    r1 = rng.normal(size=(B, 32, stepper) if per_env else (32, 7))
    r2 = rng.normal(size=r1.shape)
    off = rng.normal(size=B) * 0.3 if per_env else np.full(B, 0.2)
    coef = hm.forcing_spectrum_coefficients(r1, r2, off, L, dt, stepper, N, B)
    assert coef.shape == ((B if per_env else 1), stepper, 3, 2)
    o = BurgerOracle(B=B, L=L, N=N, dt=dt, forcing=True, stepper=stepper, offset=off)
    o.set_forcing_tables(r1, r2)
    for c in range(stepper):
        o.ioutnum = c
        F = fft(o._stochastic(), axis=-1)
        for e in range(B):
            got = coef[e if per_env else 0, c, :, 0] + 1j * coef[e if per_env else 0, c, :, 1]
            assert np.max(np.abs(got - F[e, 1:4])) < 1e-12 * np.max(np.abs(F[e]))
            rest = np.delete(F[e], [1, 2, 3, N - 1, N - 2, N - 3])
            assert np.max(np.abs(rest)) < 1e-12 * np.max(np.abs(F[e]))


def test_etdrk4_tables_match_golden(golden):
    g = golden("ks.npz")
    T = hm.etdrk4_coefficients(22.0, 64, 0.25)
    for name in ("E", "E2", "Q", "f1", "f2", "f3", "g"):
        assert np.array_equal(T[name], g["n64/" + name]), name


def test_truth_interpolant_follows_interp2d_semantics(golden):
    g = golden("burger_env.npz")
    f = hm.TruthInterpolant(g["mse_noise_A4/dns_x"], g["mse_noise_A4/dns_tt"][::10], g["mse_noise_A4/dns_uu"])
    x = hm.grid_points(2 * np.pi, 32)
    row = f(x[::-1], 0.1)                  # unsorted in, sorted out, 1-D for scalar t
    assert row.shape == (32,)
    assert np.allclose(row, f.rows(x, [0.1])[0])
    assert f(x, [0.0, 0.1]).shape == (2, 32)


def test_kernel_integer_and_division_identities():
    """Two arithmetic identities the step kernels rely on (csrc/burgers_warp.cuh, csrc/common.cuh), restated exactly:
    (1) the (agent, position) split of the state gather, floor(o m / 2^32) with m = floor(2^32 / RL) + 1, equals o // RL for
        every state layout the warp kernels accept (N <= 256, any agent count dividing N, state versions 0..4);
    (2) a / b computed as q0 + (a - q0 b) y with y = RN(1 / b), q0 = RN(a y) and fused multiply-adds is the correctly
        rounded quotient (the reward's divisions use reciprocals prepared ahead of time)."""
    from fractions import Fraction
    import random
    for N in (8, 16, 32, 64, 128, 256):
        for A in [a for a in range(1, N + 1) if N % a == 0]:
            for ver in range(5):
                nf = 2 if ver in (1, 2) else 1
                seg = N if A == 1 else N // A + 2
                RL = nf * seg + (N // 2 if ver in (3, 4) else 0)
                m = (1 << 32) // RL + 1
                assert m < (1 << 32)
                o = np.arange(A * RL + 1, dtype=np.uint64)
                assert np.array_equal((o * np.uint64(m)) >> np.uint64(32), o // np.uint64(RL)), (N, A, ver)

    def fma(x, y, z):
        return float(Fraction(x) * Fraction(y) + Fraction(z))       # one rounding

    rnd = random.Random(7)
    for _ in range(20000):
        b = rnd.choice([float(rnd.randint(1, 6000)), 15.0, rnd.uniform(1, 2) * 10 ** rnd.uniform(-12, 3)])
        a = rnd.choice([0.0, float(np.float32(rnd.uniform(0, 1) * 10 ** rnd.uniform(-12, 2))), rnd.uniform(0, 2) * 10 ** rnd.uniform(-14, 5)])
        y = 1.0 / b
        q0 = a * y
        assert fma(fma(-q0, b, a), y, q0) == a / b, (a, b)
