"""Environment functions for the batched Burgers solver.

``setup_dns_default`` and ``environment`` keep the signatures of the reference's
python/_model/burger_environment.py (:11-16 and :18-204), so ``run-vracer-burger.py`` /
``run-vracer-burger-marl.py`` can import this module instead; ``environment`` drives ONE Korali
sample ``s`` (``s["State"]``, ``s.update()``, ``s["Action"]``, ``s["Reward"]``, ``s["Termination"]``)
with a one-environment GPU batch.  ``BurgerEnvBatch`` is the vectorised form of the same episode for a
learner that wants B environments per call: one kernel launch per RL step.

Korali itself, testing-mode plotting and ``compute_Sgs`` diagnostics are out of scope.
"""
import numpy as np
import torch

from .Burger import Burger

episodeCount = 0
basis = 'hat'


def setup_dns_default(L, N, T, dt, nu, ic, forcing, seed, stepper, device=None):
    """burger_environment.py:11-16 -- DNS ground truth on the GPU (history + Ek_ktt recorded)."""
    dns = Burger(L=L, N=N, dt=dt, nu=nu, tend=T, case=ic, forcing=forcing, noise=0., seed=seed, s=stepper,
                 device=device, history=True)
    dns.simulate()
    dns.compute_Ek()
    return dns


def _truncated_v0(dns, offset, gridSize):
    """burger_environment.py:110-111: spectral hand-off DNS -> LES (literal phase factor)."""
    v0 = dns.v0.cpu().numpy() if isinstance(dns.v0, torch.Tensor) else np.asarray(dns.v0)
    v0off = v0 * np.exp(1j * 2 * np.pi * offset * dns.k)
    return np.concatenate((v0off[:((gridSize + 1) // 2)], v0off[-(gridSize - 1) // 2:])) * gridSize / dns.N


def _wrapped_truth(sgs, newx, t):
    """burger_environment.py:114-118 -- interp2d sorts its inputs, hence the two ascending pieces."""
    midx = np.argmax(newx)
    if midx == len(newx) - 1:
        return sgs.f_truth(newx, t)
    return np.concatenate((sgs.f_truth(newx[:midx + 1], t), sgs.f_truth(newx[midx + 1:], t)))


def environment(s, L, T, N, gridSize, numActions, dt, nu, episodeLength, ic, spectralReward, forcing, dforce, ssmforce,
                noise, seed, stepper, nunoise=False, version=0, ssm=False, dsm=False, dns_default=None, numAgents=1):
    """One episode for one Korali sample (burger_environment.py:18-204)."""
    global episodeCount
    assert not (ssm and dsm)
    testing = s["Custom Settings"]["Mode"] == "Testing"
    if testing:
        nu = s["Custom Settings"]["Viscosity"]
    ndns = len(dns_default)
    sidx = episodeCount % ndns
    if nunoise:
        dns = Burger(L=L, N=N, dt=dt, nu=nu, tend=T, case=ic, forcing=forcing, noise=0., seed=seed + sidx, s=stepper,
                     version=version, nunoise=nunoise, numAgents=1, history=True)
        dns.simulate()
        dns.compute_Ek()
        nu = dns.nu
    else:
        dns = dns_default[sidx]

    sgs = Burger(L=L, N=gridSize, dt=dt, nu=nu, tend=T, case=ic, forcing=forcing, dforce=dforce, ssmforce=ssmforce,
                 noise=noise, seed=seed + sidx, s=stepper, version=version, numAgents=numAgents, device=dns.device)
    sgs.randfac1 = dns.randfac1                       # :99-100
    sgs.randfac2 = dns.randfac2
    sgs.setup_basis(numActions, basis)
    if spectralReward:
        sgs.IC(v0=_truncated_v0(dns, sgs.offset, gridSize))
        sgs.set_spectrum_reference(dns)
    else:
        sgs.setGroundTruth(dns.x, dns.tt, dns.uu.cpu().numpy())
        newx = sgs.x + sgs.offset
        newx[newx > L] = newx[newx > L] - L
        newx[newx < 0] = newx[newx < 0] + L
        sgs.IC(u0=_wrapped_truth(sgs, newx, 0.))

    state = sgs.getState()
    s["State"] = state[0] if numAgents == 1 else state

    error, step = 0, 0
    nIntermediate = int(T / dt / episodeLength)       # :129 (float floor, quirk Q10)
    assert nIntermediate > 0, "dt or episodeLendth too long"
    cumreward = np.zeros(numAgents)
    reward = np.zeros(numAgents)

    while step < episodeLength and error == 0:
        s.update()                                    # policy forward in Korali
        actions = s["Action"]
        if spectralReward:
            st, rw = sgs.step_n(actions, nIntermediate)       # :148-176 as ONE launch
        else:
            sgs._ensure_truth(sgs.offset)                     # getMseReward(sgs.offset) after every sub-step, averaged
            st, rw = sgs.step_n(actions, nIntermediate)
        if int(sgs.status) != 0:
            print("[burger_environment] Exception occured during stepping:")
            error = 1
            break
        state = sgs.getState()
        if not np.isfinite(state).all():
            print("[burger_environment] Nan state detected")
            error = 1
            break
        s["State"] = state[0] if numAgents == 1 else state
        reward = rw[0].cpu().numpy().copy()
        cumreward += reward
        if not np.isfinite(reward).all():
            print("[burger_environment] Nan reward detected")
            error = 1
            break
        s["Reward"] = reward.tolist() if numAgents > 1 else reward[0]
        step += 1

    episodeCount += 1
    print(f"Episode {episodeCount}: {cumreward}")
    if error == 1:
        s["State"] = state[0] if numAgents == 1 else state
        s["Reward"] = -np.inf if numAgents == 1 else [-np.inf] * numAgents
        s["Termination"] = "Truncated"
    else:
        s["Termination"] = "Terminal"
    return sgs


class BurgerEnvBatch:
    """B copies of the episode above stepped in lock-step on one GPU.

    reset() -> states [B,S]; step(actions [B,M]) -> (states, rewards [B,A], truncated [B] bool).
    Environment e uses DNS ``e % len(dns_default)`` (the reference cycles through them per episode).
    """

    def __init__(self, B, L, T, N, gridSize, numActions, dt, nu, episodeLength, ic, spectralReward, forcing, dforce,
                 noise, seed, stepper, version=0, dns_default=None, numAgents=1, offset=None):
        self.B, self.L, self.gridSize, self.spectral = B, L, gridSize, spectralReward
        self.dns = dns_default
        self.nInt = int(T / dt / episodeLength)
        self.episodeLength = episodeLength
        ndns = len(dns_default)
        self.dmap = np.arange(B) % ndns
        d0 = dns_default[0]
        self.sgs = Burger(L=L, N=gridSize, dt=dt, nu=nu, tend=T, case='zero', forcing=forcing, dforce=dforce, noise=noise,
                          seed=seed, s=stepper, version=version, numAgents=numAgents, nenvs=B, device=d0.device,
                          history=False, offset=offset)
        if forcing:
            if ndns == 1:
                self.sgs.randfac1, self.sgs.randfac2 = d0.randfac1, d0.randfac2
            else:
                self.sgs.randfac1 = np.stack([dns_default[i].randfac1[:, :stepper] for i in self.dmap])
                self.sgs.randfac2 = np.stack([dns_default[i].randfac2[:, :stepper] for i in self.dmap])
        self.sgs.setup_basis(numActions, basis)
        if spectralReward:
            # all DNS initial spectra on the device once: reset() is ONE hand-off launch for the whole batch
            self._dns_v0 = torch.stack([torch.as_tensor(d.v0).to(device=d0.device, dtype=torch.complex128).reshape(-1)
                                        for d in dns_default])
            self._dns_k = np.asarray(d0.k, dtype=np.float64)
            ref = torch.cat([d._ektt[:, :, :gridSize // 2] for d in dns_default], dim=0)
            self.sgs.set_spectrum_reference(ref, env_map=self.dmap if ndns > 1 else None)
        else:
            off = np.broadcast_to(np.asarray(self.sgs.offset, dtype=np.float64), (B,))
            keys = np.stack([self.dmap.astype(np.float64), off], axis=1)
            uniq, inv = np.unique(keys, axis=0, return_inverse=True)
            tabs = []
            for di, sh in uniq:
                d = dns_default[int(di)]
                self.sgs.setGroundTruth(d.x, d.tt, d.uu.cpu().numpy())
                newx = self.sgs.x + sh
                newx[newx > L] -= L
                newx[newx < 0] += L
                tabs.append(self.sgs.f_truth.rows_device(newx[None], self.sgs.tt, d0.device, torch.float64)[0].cpu().numpy())
            self._truth = np.stack(tabs)
            self._tmap = inv.astype(np.int32)
            self.sgs.set_truth_table(self._truth, env_map=self._tmap if len(uniq) > 1 else None)
            self._u0_dev = torch.as_tensor(self._truth[self._tmap, 0], device=d0.device)

    def reset(self):
        B, g = self.B, self.gridSize
        off = np.broadcast_to(np.asarray(self.sgs.offset, dtype=np.float64), (B,))
        if self.spectral:          # burger_environment.py:109-112 for every environment, on the device
            self.sgs.IC_handoff(self._dns_v0, self._dns_k, src_map=self.dmap if len(self.dns) > 1 else None,
                                offsets=off if np.any(off != 0.) else None)
        else:
            self.sgs.IC(u0=self._u0_dev)
        self.step_count = 0
        return self.sgs.getState(as_tensor=True)

    def step(self, actions):
        st, rw = self.sgs.step_n(actions, self.nInt)
        self.step_count += 1
        return st, rw, self.sgs.status != 0
