"""Environment functions for the batched KS solver, with the signatures of the reference's
python/_model/ks_environment.py (``setup_dns_default`` :18-34, ``environment`` :36-120).

``environment`` drives ONE Korali sample ``s`` with a one-environment GPU batch: LES initial condition by
spectral truncation of the DNS (:52-54), ``nIntermediate`` ETDRK4 steps per action as one kernel launch,
state from the float32 row (``getState``), spectral reward against the DNS time-averaged spectrum (:98-100).
Testing-mode file output / plotting are out of scope.
"""
import numpy as np
import torch

from .KS import KS

# module-level constants of the reference (ks_environment.py:5-13)
N = 1024
L = 22
nu = 1.0
dt = 0.25
tTransient = 50
tEnd = 550
tSim = tEnd - tTransient
nSimSteps = int(tSim / dt)
basis = 'hat'


def setup_dns_default(N, dt, nu, seed, u0=None, device=None):
    """ks_environment.py:18-34: transient run, restart from its last (float32) field, main run with
    history and time-averaged spectrum.  ``u0`` pins the otherwise unseeded noise IC."""
    print("[ks_environment] setting up default dns")
    dns = KS(L=L, N=N, dt=dt, nu=nu, tend=tTransient, seed=seed, u0=u0, device=device, history=True)
    dns.simulate()
    dns.fou2real()
    u_restart = dns.uu[-1].clone()
    dns.IC(u0=u_restart)
    dns.simulate(nsteps=int(tSim / dt), restart=True)
    dns.fou2real()
    dns.compute_Ek()
    return dns


def environment(s, N, gridSize, numActions, dt, nu, episodeLength, dforce, seed, dns_default):
    """One episode for one Korali sample (ks_environment.py:36-120)."""
    testing = s["Custom Settings"]["Mode"] == "Testing"
    dns = dns_default
    v_restart = dns.vv[0]
    v_restart = v_restart.cpu().numpy() if isinstance(v_restart, torch.Tensor) else np.asarray(v_restart)
    device = getattr(dns, "device", None)

    sgs = KS(L=L, N=gridSize, dt=dt, nu=nu, tend=tSim, dforce=dforce, noise=0., device=device)
    v0 = np.concatenate((v_restart[:((gridSize + 1) // 2)], v_restart[-(gridSize - 1) // 2:])) * gridSize / dns.N
    sgs.IC(v0=v0)
    sgs.setup_basis(numActions, basis)
    sgs.set_spectrum_reference(dns if isinstance(dns, KS) else np.asarray(dns.Ek_ktt)[:, :gridSize // 2])

    s["State"] = sgs.getState().flatten().tolist()

    error, step = 0, 0
    nIntermediate = int(tSim / dt / episodeLength)
    reward, cumreward = 0., 0.
    while step < episodeLength and error == 0:
        s.update()
        actions = s["Action"]
        st, rw = sgs.step_n(actions, nIntermediate)          # :79-82 + :91 + :98-100 as one launch
        if int(sgs.status) != 0:
            print("[ks_environment] Exception occured:")
            error = 1
            break
        state = st[0].cpu().numpy().astype(np.float32).flatten().tolist()
        if np.isnan(state).any():
            print("[ks_environment] Nan state detected")
            error = 1
            break
        s["State"] = state
        reward = float(rw[0, 0])
        cumreward += reward
        if np.isnan(reward):
            print("[ks_environment] Nan reward detected")
            error = 1
            break
        s["Reward"] = reward
        step += 1
    print(cumreward)
    if error == 1:
        s["Termination"] = "Truncated"
        s["Reward"] = -1000 if testing else -np.inf
    else:
        s["Termination"] = "Terminal"
    return sgs


class KSEnvBatch:
    """B copies of the episode above (ks_environment.py:36-120) stepped in lock-step on one GPU.

    reset() -> states [B, 2 gridSize]; step(actions [B, M]) -> (states, rewards [B, 1], truncated [B] bool).
    Environment e restarts from ``dns_default[e % len(dns_default)]`` (LES initial condition by spectral truncation of
    that DNS's first history row, :52-54, as ONE hand-off launch for the whole batch) and is scored against that DNS's
    time-averaged spectrum (:98-100)."""

    def __init__(self, B, N, gridSize, numActions, dt, nu, episodeLength, dforce, seed, dns_default):
        dns_list = list(dns_default) if isinstance(dns_default, (list, tuple)) else [dns_default]
        self.B, self.gridSize, self.dns = int(B), gridSize, dns_list
        d0 = dns_list[0]
        self.nInt = int(tSim / dt / episodeLength)
        self.episodeLength = episodeLength
        self.dmap = np.arange(self.B) % len(dns_list)
        self.sgs = KS(L=L, N=gridSize, dt=dt, nu=nu, tend=tSim, dforce=dforce, noise=0., nenvs=self.B, device=d0.device,
                      history=False, u0=np.zeros(gridSize))
        self.sgs.setup_basis(numActions, basis)
        # first history row of every DNS (complex64, as the reference reads dns.vv[0]) on the device once
        self._dns_v0 = torch.stack([torch.as_tensor(d.vv[0]).to(device=d0.device, dtype=torch.complex128).reshape(-1)
                                    for d in dns_list])
        self._dns_k = np.asarray(d0.k, dtype=np.float64)
        ref = torch.cat([d._ektt[:, :, :gridSize // 2] for d in dns_list], dim=0)
        self.sgs.set_spectrum_reference(ref, env_map=self.dmap if len(dns_list) > 1 else None)
        self.step_count = 0

    def reset(self):
        self.sgs.IC_handoff(self._dns_v0, self._dns_k, src_map=self.dmap if len(self.dns) > 1 else None)
        self.step_count = 0
        return self.sgs.getState(as_tensor=True)

    def step(self, actions):
        st, rw = self.sgs.step_n(actions, self.nInt)
        self.step_count += 1
        return st, rw, self.sgs.status != 0
