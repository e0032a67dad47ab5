"""Batched, GPU-resident drop-in for the reference ``KS`` (Kuramoto-Sivashinsky) class.

Mirrors /root/reference/python/_model/KS.py: constructor keywords :33, ``setup_basis`` :139,
``IC`` :166, ``setGroundTruth`` :221 (argument order t, x, uu), ``step`` :230, ``simulate`` :276,
``fou2real`` :316, ``compute_Ek`` :322, ``getState`` :369.  ETDRK4 tables are computed on the
host exactly as the reference does (numpy, 62-point contour means) and uploaded once; every
``step`` runs in the CUDA library.  ``nenvs == 1`` returns what the reference returns.
"""
import numpy as np
import torch

from . import _lib as LB
from ._spectral import SpectralEnv
from .hostmath import grid_points, etdrk4_coefficients, TruthInterpolant

L_check = LB.check


class KS(SpectralEnv):
    equation = LB.KS

    def __init__(self, L=2. * np.pi, N=512, dt=0.001, nu=1.0, dforce=True, nsteps=None, tend=5., u0=None, v0=None,
                 case=None, noise=0., seed=42, *, nenvs=1, device=None, dtype=torch.float64, history=None,
                 ic_seed=None):
        B = int(nenvs)
        self.noise, self.seed = noise, seed
        self.L, self.dt, self.tend = float(L), dt, float(tend)
        nsteps = int(tend / dt) if nsteps is None else int(nsteps)
        self.N, self.dx = N, L / N
        self.x = grid_points(self.L, N)
        self.nu = nu                                  # stored but unused by the reference's operator (KS.py:117)
        self.nsteps = self.nout = nsteps
        self.sigma = L / (2 * N)
        self.M, self.basis, self.f_truth, self.uu_truth = 0, None, None, None
        self.dforce = bool(dforce)
        self.numAgents = 1
        self._ic_rng = np.random.default_rng(ic_seed)  # the reference draws its noise IC unseeded (KS.py:36,175)

        self._create(nenvs=B, N=N, L_=self.L, dt=float(dt), M=0, num_agents=1, version=0, stepper=1,
                     flags=(LB.DFORCE if dforce else 0), device=device, dtype=dtype)
        tab = etdrk4_coefficients(self.L, N, dt)      # KS.py:112-137
        self.k, self.l = tab['k'], tab['l']
        for name in ('E', 'E2', 'Q', 'f1', 'f2', 'f3', 'g'):
            setattr(self, name, tab[name])
        args = [LB.as_dp(np.ascontiguousarray(tab[n], dtype=np.float64)) for n in ('E', 'E2', 'Q', 'f1', 'f2', 'f3')]
        L_check(self._lib.mpde_set_etdrk4(self._h, *args))

        self._setup_history(history)
        self._state_buf = torch.empty((B, self._state_size), device=self.device, dtype=self.dtype)
        self._reward_buf = torch.zeros((B, 1), device=self.device, dtype=self.dtype)
        self._state_at = -1
        self._uu_valid_at = -1
        self._spec_ref = None

        if case is not None:
            self.IC(case=case)
        elif u0 is None and v0 is None:
            self.IC()
        elif u0 is not None:
            self.IC(u0=u0)
        else:
            self.IC(v0=v0)

    # ------------------------------------------------------------------ set-up
    def IC(self, u0=None, v0=None, case='noise', seed=42, mask=None):
        """KS.py:166-219."""
        N = self.N
        if v0 is None:
            if u0 is None:
                if case != 'noise':
                    print("[KS] Error: IC case unknown")
                    return -1
                u0 = self._ic_rng.normal(0., 1e-3, (self.nenvs, N))
            elif np.shape(u0)[-1] != N:
                raise SystemExit("[KS] Error: wrong IC array size")
            u0d = self._batch(u0, self.dtype, (N,))
            m, mp = self._mask_ptr(mask)
            L_check(self._lib.mpde_reset_u(self._h, self._ptr(u0d), mp, self._stream()))
        else:
            if np.shape(v0)[-1] != N:
                raise SystemExit("[KS] Error: wrong IC array size")
            v0d = self._batch(v0, self.cdtype, (N,))
            m, mp = self._mask_ptr(mask)
            L_check(self._lib.mpde_reset_v(self._h, self._ptr(torch.view_as_real(v0d)), mp, self._stream()))
        self.t = 0.
        self.stepnum = 0
        self.ioutnum = 0
        self._state_at = -1
        self._uu_valid_at = -1
        self.u0 = self.u
        self.v0 = self.v

    def setGroundTruth(self, t, x, uu):
        """KS.py:221-223 (argument order t, x, uu)."""
        self.uu_truth = uu
        self.f_truth = TruthInterpolant(_np(x), _np(t), _np(uu), kind='cubic')

    def mapGroundTruth(self):
        t = np.arange(0, self.nout + 1) * self.dt
        return self.f_truth(self.x, t)

    # ------------------------------------------------------------------ stepping
    def _actions(self, actions):
        if actions is None:
            return None
        assert self.basis is not None, "[KS] Basis not set up (is None)."
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=self.dtype)
        else:
            a = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self.device).to(self.dtype)
        a = a.reshape(self.nenvs, -1)
        assert a.shape[1] == self.M, "[KS] Wrong number of actions (provided {}/ expected {})".format(a.shape[1], self.M)
        return a.contiguous()

    def step_n(self, actions=None, n=1, want_state=True, want_reward=True):
        """``n`` ETDRK4 steps with the same actions (ks_environment.py:79-80) + getState + spectral
        reward in one launch.  Returns (state [B, 2N] or None, reward [B, 1] or None)."""
        a = self._actions(actions)
        if a is not None and not self.dforce:
            if self._uu_valid_at < 0:
                # the reference crashes here: uu is still complex64 before the first fou2real (KS.py:241-245)
                raise TypeError("[KS] dforce=False needs fou2real()/getState() before the first step")
            L_check(self._lib.mpde_set_option(self._h, LB.OPT_KS_UUROW, int(self._uu_valid_at == self.ioutnum)))
        st = self._state_buf if want_state else None
        rw = self._reward_buf if (want_reward and self._spec_ref is not None) else None
        L_check(self._lib.mpde_step(self._h, self._ptr(a), int(n), self._ptr(st), self._ptr(rw), self._stream()))
        self.stepnum += n
        self.ioutnum += n
        for _ in range(n):
            self.t += self.dt
        if st is not None:
            self._state_at = self.ioutnum
            self._uu_valid_at = self.ioutnum          # getState ran fou2real
        return st, rw

    def step(self, actions=None):
        """KS.py:230-274: one ETDRK4 step."""
        self.step_n(actions, 1, want_state=False, want_reward=False)

    def simulate(self, nsteps=None, restart=False, correction=[]):
        """KS.py:276-314 (``correction`` is not supported)."""
        if len(correction):
            raise NotImplementedError("simulate(correction=...)")
        if nsteps is not None:
            self.nsteps = int(nsteps)
        if restart:
            self.nout = self.nsteps
            self._setup_history(self.history)
            self.IC(v0=self.v0 if self.nenvs > 1 else self.v0[None])
        left = self.nsteps
        while left > 0:
            n = min(left, 200)
            self.step_n(None, n, want_state=False, want_reward=False)
            left -= n
        if bool((self.status != 0).any()):
            print("[KS] Floating point exception occured", flush=True)
            return -1

    # ------------------------------------------------------------------ observables
    def fou2real(self):
        """KS.py:316-320: the float32 real-space history exists as ``uu`` (written by the kernel);
        this only marks row ``ioutnum`` as current for a following dforce=False step."""
        self._uu_valid_at = self.ioutnum

    def getState(self, as_tensor=None):
        """KS.py:369-383 -> [dudx ; d2udx2] evaluated on the float32 row (numpy [2N] for nenvs == 1)."""
        if self._state_at != self.ioutnum:
            L_check(self._lib.mpde_step(self._h, None, 0, self._ptr(self._state_buf), None, self._stream()))
            self._state_at = self.ioutnum
        self._uu_valid_at = self.ioutnum
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return self._state_buf if as_tensor else self._state_buf[0].cpu().numpy().astype(np.float32)

    def getReward(self):
        """KS.py:360-367: -|u - truth(x, t)| per grid point (float32 u row)."""
        u = self.uu[..., self.ioutnum, :] if self.history else self.u
        truth = torch.as_tensor(self.f_truth(self.x, [self.t]), device=self.device)
        r = -(u - truth).abs()
        return r if self.nenvs > 1 else r.cpu().numpy()


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
