"""Environment-level parity: marlpde_b200.burger_environment.environment (same signature as the
reference's python/_model/burger_environment.py) driven by a fake Korali sample reproduces the states
and rewards the REFERENCE environment function produced for the same scripted actions
(tests/golden/burger_env.npz), with the DNS ground truth computed on the GPU as well."""
import functools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TWO_PI = 2 * np.pi


class FakeSample(dict):
    def __init__(self, script):
        super().__init__()
        self["Custom Settings"] = {"Mode": "Training"}
        self.script, self.i = script, 0
        self.states, self.rewards = [], []

    def update(self):
        if self.i == 0:
            self.state0 = np.array(self["State"], dtype=float)
        else:
            self.states.append(np.array(self["State"], dtype=float))
            self.rewards.append(np.array(self["Reward"], dtype=float))
        self["Action"] = self.script[self.i]
        self.i += 1


CASES = ["spec_A1", "spec_A4", "spec_A32_v1", "spec_noise", "mse_A1", "mse_A32", "mse_noise_A4"]


@pytest.mark.parametrize("tag", CASES)
def test_environment_function_matches_reference_episode(golden, tag, monkeypatch):
    import marlpde_b200.burger_environment as be
    from marlpde_b200 import Burger
    g = golden("burger_env.npz")
    p = tag + "/"
    spectral, A, noise, forcing, dforce, ver, stepper, epl, NDNS = g[p + "cfg"]
    A, ver, stepper, epl, NDNS = int(A), int(ver), int(stepper), int(epl), int(NDNS)
    L, T, dt, nu, gsz = TWO_PI, 0.4, 1e-3, 0.02, 32
    dns = be.setup_dns_default(L, NDNS, T, dt, nu, "turbulence", bool(forcing), 50, stepper)
    # the GPU DNS itself reproduces the reference DNS
    np.testing.assert_allclose(dns.Ek_ktt.cpu().numpy()[:, :gsz // 2], g[p + "dns_Ek_ktt"], rtol=2e-5, atol=1e-25)
    if noise > 0:      # the reference draws the offset from an unseeded generator: pin the recorded one
        monkeypatch.setattr(be, "Burger", functools.partial(Burger, offset=float(g[p + "offset"])))
    acts = g[p + "actions"]
    script = [a.tolist() for a in acts] if A == 1 else [a.reshape(A, -1).tolist() for a in acts]
    s = FakeSample(script)
    be.episodeCount = 0
    sgs = be.environment(s, L, T, NDNS, gsz, 32, dt, nu, epl, "turbulence", bool(spectral), bool(forcing), bool(dforce),
                         False, noise, 50, stepper, version=ver, dns_default=[dns], numAgents=A)
    assert s["Termination"] == "Terminal"
    states = s.states + [np.array(s["State"], dtype=float)]
    rewards = s.rewards + [np.array(s["Reward"], dtype=float)]

    def rel(a, b):
        return np.max(np.abs(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1))) / np.max(np.abs(b))

    assert rel(s.state0, g[p + "state0"]) < 1e-9
    for i in range(epl):
        assert rel(states[i], g[p + "states"][i]) < 1e-9, (tag, i)
        np.testing.assert_allclose(np.atleast_1d(rewards[i]), np.atleast_1d(g[p + "rewards"][i]), rtol=2e-5, atol=1e-8,
                                   err_msg=f"{tag} step {i}")
    assert rel(sgs.u.cpu().numpy(), g[p + "sgs_u_final"]) < 1e-9


def test_batched_episode_equals_single_sample_episodes(golden):
    """BurgerEnvBatch (B envs per launch) == B independent environment() episodes, bitwise."""
    import marlpde_b200.burger_environment as be
    L, T, dt, nu, gsz, epl = TWO_PI, 0.2, 1e-3, 0.02, 32, 20
    dns = [be.setup_dns_default(L, 256, T, dt, nu, "turbulence", True, 50 + i * 9, 1) for i in range(2)]
    B = 4
    rng = np.random.default_rng(5)
    acts = rng.uniform(0.0, 0.03, (epl, B, 32))
    batch = be.BurgerEnvBatch(B, L, T, 256, gsz, 32, dt, nu, epl, "turbulence", True, True, False, 0., 50, 1,
                              dns_default=dns)
    st = [batch.reset().clone()]
    rws = []
    for i in range(epl):
        s_, r_, trunc = batch.step(acts[i])
        st.append(s_.clone()); rws.append(r_.clone())
        assert not bool(trunc.any())
    for e in range(B):
        s = FakeSample([a.tolist() for a in acts[:, e]])
        be.episodeCount = e            # environment() picks dns_default[episodeCount % ndns]
        be.environment(s, L, T, 256, gsz, 32, dt, nu, epl, "turbulence", True, True, False, False, 0., 50, 1,
                       dns_default=dns, numAgents=1)
        states = [s.state0] + s.states + [np.array(s["State"], dtype=float)]
        rewards = s.rewards + [np.array(s["Reward"], dtype=float)]
        for i in range(epl + 1):
            assert np.array_equal(states[i], st[i][e].cpu().numpy()), (e, i)
        for i in range(epl):
            assert float(rewards[i]) == float(rws[i][e, 0]), (e, i)


@pytest.mark.parametrize("tag,dforce", [("direct", True), ("eddy", False)])
def test_ks_environment_function_matches_reference_episode(golden, tag, dforce):
    """marlpde_b200.ks_environment.environment against the reference's ks_environment.environment
    (fake Korali sample, recorded DNS spectrum; the DNS itself is chaotic and not comparable run to run)."""
    import types
    import marlpde_b200.ks_environment as ke
    g = golden("ks_env.npz")
    NDNS, dt, gsz, epl, M = g["cfg"]
    NDNS, gsz, epl, M = int(NDNS), int(gsz), int(epl), int(M)
    dns = types.SimpleNamespace(N=NDNS, vv=g["dns_vv0"][None], Ek_ktt=g["dns_Ek_ktt"])
    acts = g[tag + "/actions"]
    s = FakeSample([a.tolist() for a in acts])
    ke.environment(s, NDNS, gsz, M, dt, 1.0, epl, dforce, 42, dns)
    assert s["Termination"] == "Terminal"
    states = np.array(s.states + [np.array(s["State"], dtype=float)])
    rewards = np.array(s.rewards + [np.array(s["Reward"], dtype=float)])
    dx = 22 / gsz
    noise = 16 * np.finfo(np.float32).eps * 4.0 / dx ** 2         # float32 field, second difference (quirk Q7)
    assert np.max(np.abs(s.state0 - g[tag + "/state0"])) <= noise
    # 40 chaotic ETDRK4 steps per action amplify round-off: compare the early episode tightly, all of it loosely
    assert np.max(np.abs(states[:5] - g[tag + "/states"][:5])) <= 50 * noise
    np.testing.assert_allclose(rewards[:5], g[tag + "/rewards"][:5], rtol=1e-3, atol=1e-7)
    if dforce:     # (the float32-row forcing of dforce=False seeds 1e-7 differences that the chaotic dynamics amplify)
        assert np.corrcoef(rewards, g[tag + "/rewards"])[0, 1] > 0.99
    else:
        np.testing.assert_allclose(rewards[:10], g[tag + "/rewards"][:10], rtol=5e-2, atol=1e-6)


def test_ks_dns_setup_runs_and_has_the_reference_spectrum(golden):
    """setup_dns_default (transient + restart + main run) from the recorded noise IC: the trajectory is chaotic,
    but the time-averaged spectrum of the attractor matches the reference DNS."""
    import marlpde_b200.ks_environment as ke
    g = golden("ks_env.npz")
    dns = ke.setup_dns_default(256, 0.25, 1.0, 42, u0=g["transient_u0"])
    assert dns.ioutnum == 2000 and int(dns.status) == 0
    ek = dns.Ek_ktt.cpu().numpy()[-1, 1:9]
    np.testing.assert_allclose(ek, g["dns_Ek_ktt"][-1, 1:9], rtol=0.35)


def test_ks_batched_episode_equals_single_sample_episodes(golden):
    """KSEnvBatch (B environments in lock-step, device-side hand-off) == B runs of ks_environment.environment, bitwise."""
    import marlpde_b200.ks_environment as ke
    g = golden("ks_env.npz")
    dns = ke.setup_dns_default(256, 0.25, 1.0, 42, u0=g["transient_u0"])
    gsz, M, epl, B = 32, 16, 50, 3
    rng = np.random.default_rng(4)
    acts = rng.normal(0.0, 0.05, (B, epl, M))
    env = ke.KSEnvBatch(B, 256, gsz, M, 0.25, 1.0, epl, True, 42, dns)
    st0 = env.reset().cpu().numpy()
    S, R = [], []
    for i in range(6):
        st, rw, trunc = env.step(torch.as_tensor(acts[:, i], device=env.sgs.device))
        assert not bool(trunc.any())
        S.append(st.cpu().numpy().copy()); R.append(rw.cpu().numpy().copy())
    for e in range(B):
        s = FakeSample([a.tolist() for a in acts[e, :6]] + [acts[e, 6].tolist()] * (epl - 6))
        sgs = ke.environment(s, 256, gsz, M, 0.25, 1.0, epl, True, 42, dns)
        assert np.array_equal(np.asarray(s.state0, dtype=np.float32), st0[e].astype(np.float32))
        for i in range(6):
            assert np.array_equal(np.asarray(s.states[i], dtype=np.float32), S[i][e].astype(np.float32)), (e, i)
            assert float(s.rewards[i]) == float(R[i][e, 0]), (e, i)
