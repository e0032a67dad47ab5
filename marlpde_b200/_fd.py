"""Batched drop-ins for the reference finite-difference environments ``Diffusion`` and
``Advection`` (/root/reference/python/_model/Diffusion.py, Advection.py).

In the reference the agents' actions are the entries of a periodic 3-point stencil that is
applied as a dense N x N matrix product each step (Diffusion.py:164-206, Advection.py:154-200);
here the stencil is applied directly by a CUDA kernel, one warp per environment.  ``numAgents``
is a per-call argument exactly as in the reference.  ``nenvs == 1`` returns python lists /
floats like the reference, ``nenvs > 1`` device tensors with the environment index first.
"""
import numpy as np
import torch

from . import _lib as LB
from ._base import BatchedEnv
from .hostmath import grid_points, TruthInterpolant

L_check = LB.check


class _FDEnv(BatchedEnv):
    name = "FD"

    def _common_init(self, L, N, dt, nu, nsteps, tend, case, version, noise, nunoise, seed, implicit, nenvs, device, dtype,
                     history, offset, round_steps):
        B = int(nenvs)
        self.seed, self.implicit = seed, bool(implicit)
        self.L, self.dt, self.tend = float(L), float(dt), float(tend)
        if nsteps is None:
            nsteps = int(tend / dt + 0.5) if round_steps else int(tend / dt)
        else:
            nsteps = int(nsteps)
            self.tend = dt * nsteps
        self.N, self.dx = N, L / N
        self.x = grid_points(self.L, N)
        rng = np.random.default_rng()                 # the reference draws these from an unseeded generator
        self._nu0 = float(nu)
        self._nu = np.full(B, float(nu))
        if nunoise:
            self._nu = 0.01 + 0.02 * rng.uniform(size=B)
        self.noise = noise
        if offset is not None:
            self._offset = np.broadcast_to(np.asarray(offset, dtype=np.float64), (B,)).copy()
        else:
            self._offset = rng.normal(0., noise, B) if noise > 0. else np.zeros(B)
        self.nsteps = self.nout = nsteps
        self.version = version
        if version > 1:
            raise SystemExit(f"[{self.name}] Version not recognized")
        self.f_truth, self.uu_truth = None, None
        self.case = case
        self._create(nenvs=B, N=N, L_=self.L, dt=self.dt, M=0, num_agents=1, version=0, stepper=1,
                     flags=(LB.IMPLICIT if implicit else 0), device=device, dtype=dtype)
        L_check(self._lib.mpde_set_nu(self._h, LB.as_dp(np.ascontiguousarray(self._nu)), B))
        rows = self.nout + 1
        if history is None:
            history = B * rows * N * 8 <= (2 << 30)
        self.history = bool(history)
        self.tt = np.concatenate(([0.], np.cumsum(np.full(self.nout, self.dt))))
        if self.history:
            self._uu = torch.zeros((B, rows, N), device=self.device, dtype=self.dtype)
            L_check(self._lib.mpde_set_history(self._h, self._ptr(self._uu), None, None, rows))
        self._A, self._M = 1, 0
        self._truth_key = None
        if case is None:
            raise SystemExit(f"[{self.name}] IC ambigous")
        self.IC(case=case)

    # ------------------------------------------------------------------ attributes
    @property
    def nu(self):
        return float(self._nu[0]) if self.nenvs == 1 else self._nu

    @property
    def offset(self):
        return float(self._offset[0]) if self.nenvs == 1 else self._offset

    def _squeeze(self, t):
        return t[0] if self.nenvs == 1 else t

    @property
    def u(self):
        return self._squeeze(self._get(LB.FIELD_U, (self.nenvs, self.N), self.dtype))

    @property
    def uu(self):
        if not self.history:
            raise RuntimeError("history recording is off for this batch (pass history=True)")
        return self._squeeze(self._uu)

    def _set_agents(self, A):
        if A != self._A:
            assert self.N % A == 0
            L_check(self._lib.mpde_set_option(self._h, LB.OPT_NUM_AGENTS, int(A)))
            self._A = A

    def _set_actions(self, M):
        if M != self._M:
            L_check(self._lib.mpde_set_option(self._h, LB.OPT_NUM_ACTIONS, int(M)))
            self._M = M

    def _reset(self, u0):
        self._u0 = np.asarray(u0, dtype=np.float64)
        u0d = self._batch(u0, self.dtype, (self.N,))
        L_check(self._lib.mpde_reset_u(self._h, self._ptr(u0d), None, self._stream()))
        self.u0 = self._squeeze(u0d)
        self.t = 0.
        self.stepnum = 0
        self.ioutnum = 0
        self._truth_key = None

    def setGroundTruth(self, t, x, uu):
        """Diffusion.py:130-132 / Advection.py:131-133 (argument order t, x, uu; linear interpolation)."""
        self.uu_truth = uu
        self.f_truth = TruthInterpolant(_np(x), _np(t), _np(uu), kind='linear')
        self._truth_key = None

    def mapGroundTruth(self):
        return self.f_truth(self.x, self.tt)

    # ------------------------------------------------------------------ stepping
    def _flat_actions(self, actions, numAgents, per_point):
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=self.dtype)
        else:
            a = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self.device).to(self.dtype)
        a = a.reshape(self.nenvs, -1).contiguous()        # MARL lists [A][P] flatten to the grid order
        return a

    def step_n(self, actions=None, numAgents=1, n=1, want_state=False, want_reward=False):
        a = None
        if actions is not None:
            a = self._flat_actions(actions, numAgents, True)
            self._check_action_len(a.shape[1], numAgents)
            self._set_actions(a.shape[1])
        self._set_agents(numAgents)
        st = rw = None
        if want_state:
            S = self.N if numAgents == 1 else numAgents * (self.N // numAgents + 2)
            st = torch.empty((self.nenvs, S), device=self.device, dtype=self.dtype)
        if want_reward:
            self._ensure_truth()
            rw = torch.empty((self.nenvs, numAgents), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, self._ptr(a), int(n), self._ptr(st), self._ptr(rw), self._stream()))
        self.stepnum += n
        self.ioutnum += n
        for _ in range(n):
            self.t += self.dt
        return st, rw

    def step(self, actions=None, numAgents=1):
        self.step_n(actions, numAgents, 1)

    def simulate(self, nsteps=None):
        """Diffusion.py:219-236 / Advection.py:216-233."""
        left = self.nsteps - self.stepnum
        while left > 0:
            n = min(left, 500)
            self.step_n(None, 1, n)
            left -= n
        if bool((self.status != 0).any()):
            print(f"[{self.name}] Floating point exception occured in simulate", flush=True)
            return -1

    # ------------------------------------------------------------------ observables
    def getState(self, numAgents=1, as_tensor=None):
        self._set_agents(numAgents)
        S = self.N if numAgents == 1 else numAgents * (self.N // numAgents + 2)
        st = torch.empty((self.nenvs, S), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, None, 0, self._ptr(st), None, self._stream()))
        if numAgents > 1:
            st = st.view(self.nenvs, numAgents, -1)
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return st if as_tensor else st[0].cpu().numpy().tolist()

    def _truth_rows(self):
        raise NotImplementedError

    def _ensure_truth(self, offset=0.):
        key = ("truth", float(np.sum(offset)))
        if self._truth_key == key:
            return
        tabs, inv = self._truth_rows(offset)
        tab = self._dev(tabs)
        mp = None
        if inv is not None:
            self._keep['truth_map'] = self._dev(inv, torch.int32, (self.nenvs,))
            mp = self._ptr(self._keep['truth_map'])
        self._keep['truth'] = tab.contiguous()
        L_check(self._lib.mpde_set_truth(self._h, self._ptr(self._keep['truth']), tab.shape[0], tab.shape[1], mp))
        L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_MSE))
        self._truth_key = key

    def getMseReward(self, numAgents=1, offset=0., as_tensor=None):
        """Diffusion.py:238-273 / Advection.py:235-270: -mean((truth - u)^2) per agent section."""
        assert self.N % numAgents == 0
        self._ensure_truth(offset)
        self._set_agents(numAgents)
        rw = torch.empty((self.nenvs, numAgents), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, None, 0, None, self._ptr(rw), self._stream()))
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        if as_tensor:
            return rw
        r = rw[0].cpu().numpy()
        return float(r[0]) if numAgents == 1 else r.tolist()

    def _unique_rows(self, make_row_table, keys):
        uniq, inv = np.unique(np.asarray(keys), axis=0, return_inverse=True)
        tabs = np.stack([make_row_table(*k) for k in uniq])
        return tabs, (inv.astype(np.int32) if len(uniq) > 1 else None)


class Diffusion(_FDEnv):
    """u_t = nu u_xx, explicit / implicit Euler, central differences (Diffusion.py:8-307)."""
    equation = LB.DIFFUSION
    name = "Diffusion"

    def __init__(self, L=2. * np.pi, N=512, dt=0.001, nu=0.01, nsteps=None, tend=5., case='box', version=0, noise=0.,
                 nunoise=False, seed=1337, implicit=False, *, nenvs=1, device=None, dtype=torch.float64, history=None,
                 offset=None):
        self._common_init(L, N, dt, nu, nsteps, tend, case, version, noise, nunoise, seed, implicit, nenvs, device, dtype,
                          history, offset, round_steps=False)
        if not implicit and np.any(2. * self._nu * self.dt >= self.dx ** 2):
            print(f"[Diffusion] Warning: CFL condition violated {2. * self._nu.max() * self.dt}>{self.dx ** 2}", flush=True)

    def IC(self, case='box'):
        """Diffusion.py:98-128."""
        x, L = self.x[None, :], self.L
        off = self._offset[:, None]
        if case == 'box':
            u0 = np.zeros((self.nenvs, self.N))
            u0[np.abs(x - L / 2 - off) < L / 8] = 1.
        elif case == 'sinus':
            u0 = np.sin((x - off) * 2 * np.pi / L)
        elif case == 'gaussian':
            u0 = np.exp(-0.5 * (0.5 * L + off - x) ** 2)
        else:
            raise SystemExit("[Diffusion] Error: IC case unknown")
        self._reset(u0)

    def _check_action_len(self, m, numAgents):
        if numAgents == 1 and m != self.N:
            assert m == 1, f"[Diffusion] action len not 1, it is {m}"
        else:
            assert m == self.N, f"[Diffusion] need N actions in total, got {m}"

    def getAnalyticalSolution(self, t):
        """Diffusion.py:301-306."""
        if self.case == "sinus":
            sol = self._u0 * np.exp(-(2. * np.pi / self.L) ** 2 * self._nu[:, None] * t)
            return sol[0] if self.nenvs == 1 else sol
        print(f"[Diffusion] case {self.case} not available")

    @property
    def solution(self):
        """Analytic rows stored by step() for the sinus case (Diffusion.py:215-216)."""
        sol = self._u0[:, None, :] * np.exp(-(2. * np.pi / self.L) ** 2 * self._nu[:, None, None] * self.tt[None, :, None])
        return sol[0] if self.nenvs == 1 else sol

    def _truth_rows(self, offset):
        if self.case == "sinus":
            sol = self._u0[:, None, :] * np.exp(-(2. * np.pi / self.L) ** 2 * self._nu[:, None, None] * self.tt[None, :, None])
            if self.nenvs == 1 or (np.all(self._u0 == self._u0[0]) and np.all(self._nu == self._nu[0])):
                return sol[:1], None
            return sol, np.arange(self.nenvs, dtype=np.int32)
        newx = self.x - offset
        newx[newx > self.L] -= self.L
        newx[newx < 0] += self.L
        return self.f_truth.rows(newx, self.tt)[None], None

    def getDirectReward(self, numAgents=1, as_tensor=None):
        """Diffusion.py:275-281."""
        assert numAgents == self.N, f"[Diffusion] direct reward neeeds N agents (using {numAgents})"
        self._set_agents(numAgents)
        L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_DIRECT))
        rw = torch.empty((self.nenvs, self.N), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, None, 0, None, self._ptr(rw), self._stream()))
        if self._truth_key is not None:
            L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_MSE))
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return rw if as_tensor else rw[0].cpu().numpy().tolist()


class Advection(_FDEnv):
    """u_t + nu u_x = 0, Lax scheme / action stencil (Advection.py:8-295)."""
    equation = LB.ADVECTION
    name = "Advection"

    def __init__(self, L=2. * np.pi, N=512, dt=0.001, nu=0.01, nsteps=None, tend=5., case='sinus', version=0, noise=0.,
                 nunoise=False, seed=1337, implicit=False, *, nenvs=1, device=None, dtype=torch.float64, history=None,
                 offset=None):
        self._common_init(L, N, dt, nu, nsteps, tend, case, version, noise, nunoise, seed, implicit, nenvs, device, dtype,
                          history, offset, round_steps=True)
        # Courant number from the nu given to the constructor, BEFORE nunoise replaces it (Advection.py:43-46)
        self.alpha = self._nu0 * self.dt / self.dx
        al = torch.full((self.nenvs,), self.alpha, device=self.device, dtype=self.dtype)
        self._set(LB.FIELD_ALPHA, al)
        if np.any(self._nu > self.dx / self.dt):
            print("[Advection] Warning: CFL condition violated", flush=True)

    def IC(self, case='box'):
        """Advection.py:97-129 (only 'sinus' is implemented by the reference)."""
        if case != 'sinus':
            assert False, "Not yet implemented"
        self._reset(np.sin((self.x[None, :] - self._offset[:, None]) * 2 * np.pi / self.L))

    def _check_action_len(self, m, numAgents):
        if numAgents == 1 and m != 2 * self.N:
            assert m == 2, f"[Advection] action len not 1, it is {m}"
        else:
            assert m == 2 * self.N, f"[Advection] need 2N actions in total, got {m}"

    def getAnalyticalSolution(self, t):
        """Advection.py:289-294."""
        sol = np.sin((self.x[None, :] - self._nu[:, None] * t - self._offset[:, None]) * 2 * np.pi / self.L)
        return sol[0] if self.nenvs == 1 else sol

    @property
    def solution(self):
        sol = np.sin((self.x[None, None, :] - self._nu[:, None, None] * self.tt[None, :, None] - self._offset[:, None, None])
                     * 2 * np.pi / self.L)
        sol[:, 0] = self._u0
        return sol[0] if self.nenvs == 1 else sol

    def _truth_rows(self, offset):
        sol = np.sin((self.x[None, None, :] - self._nu[:, None, None] * self.tt[None, :, None] - self._offset[:, None, None])
                     * 2 * np.pi / self.L)
        if self.nenvs == 1 or (np.all(self._nu == self._nu[0]) and np.all(self._offset == self._offset[0])):
            return sol[:1], None
        return sol, np.arange(self.nenvs, dtype=np.int32)


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class DiffusionError(Diffusion):
    """Diffusion whose actions are the ERROR of the Laplacian stencil: row k of the update matrix is
    (1 - a_k/2, -2 + a_k, 1 - a_k/2) (/root/reference/python/_model/DiffusionError.py:160-216).  Same constructor, IC cases,
    state and rewards as ``Diffusion``; only ``step(actions, numAgents)`` differs (in the kernel: fd_launch.cu)."""
    equation = LB.DIFFUSION_ERROR
    name = "Diffusion"          # the reference keeps the [Diffusion] prefix in its messages


class Laplace(BatchedEnv):
    """Relaxation towards u_xx = f on N + 1 points with one Dirichlet point u[0] = 1; agent i owns row i + 1 of the update
    matrix and sets its three entries (/root/reference/python/_model/Laplace.py:8-166).  ``step(actions, numAgents)`` with
    ``numAgents == N`` (the class's N, i.e. constructor N + 1 points minus the boundary point) and three actions per agent;
    ``getState`` -> [u_{i-1}, u_i, u_{i+1}, force_i] per agent, ``getDirectReward`` -> -(u_xx - force)^2 on points 1..N."""
    equation = LB.LAPLACE

    def __init__(self, L=2. * np.pi, N=512, dt=0.01, ic='one', sforce='zero', episodeLength=100, noise=0., version=0, *,
                 nenvs=1, device=None, dtype=torch.float64, history=None, offset=None, force_seed=None):
        B = int(nenvs)
        self.N = int(N) + 1                                              # Laplace.py:18
        self.L, self.dt = float(L), float(dt)
        self.nsteps = self.nout = int(episodeLength)
        rng = np.random.default_rng(force_seed)                         # the reference draws from the unseeded global stream
        self._rng = rng
        if offset is not None:
            self._offset = np.broadcast_to(np.asarray(offset, dtype=np.float64), (B,)).copy()
        else:
            self._offset = rng.normal(0., self.L * noise, B) if noise > 0. else np.zeros(B)
        self.version = version
        self.dx = L / self.N
        self.x = np.linspace(0, self.L, self.N, endpoint=False)
        self._create(nenvs=B, N=self.N, L_=self.L, dt=self.dt, M=3 * (self.N - 1), num_agents=1, version=0, stepper=1,
                     flags=0, reward_mode=LB.REWARD_DIRECT, device=device, dtype=dtype)
        L_check(self._lib.mpde_set_nu(self._h, LB.as_dp(np.zeros(B)), B))
        rows = self.nout + 1
        if history is None:
            history = B * rows * self.N * 8 <= (2 << 30)
        self.history = bool(history)
        self.tt = np.concatenate(([0.], np.cumsum(np.full(self.nout, self.dt))))
        if self.history:
            self._uu = torch.zeros((B, rows, self.N), device=self.device, dtype=self.dtype)
            L_check(self._lib.mpde_set_history(self._h, self._ptr(self._uu), None, None, rows))
        self.ic, self.sforce = ic, sforce
        self.IC(ic=ic, sforce=sforce)

    @property
    def offset(self):
        return float(self._offset[0]) if self.nenvs == 1 else self._offset

    def _squeeze(self, t):
        return t[0] if self.nenvs == 1 else t

    @property
    def u(self):
        return self._squeeze(self._get(LB.FIELD_U, (self.nenvs, self.N), self.dtype))

    @property
    def uu(self):
        if not self.history:
            raise RuntimeError("history recording is off for this batch (pass history=True)")
        return self._squeeze(self._uu)

    def IC(self, ic='one', sforce='zero'):
        """Laplace.py:46-114."""
        B, x, L = self.nenvs, self.x[None, :], self.L
        off = self._offset[:, None]
        if ic == 'zero':
            u0 = np.zeros((B, self.N))
        elif ic == 'one':
            u0 = np.ones((B, self.N))
        elif ic == 'sin':
            u0 = np.broadcast_to(1. + np.sin(x), (B, self.N)).copy()
        elif ic == 'cos':
            u0 = np.broadcast_to(np.cos(x), (B, self.N)).copy()
        else:
            raise SystemExit("[Laplace] Error: ic unknown")
        sin1, cos1 = np.sin((x - off) * 2 * np.pi / L), np.cos((x - off) * 2 * np.pi / L)
        if sforce == 'zero':
            force = np.zeros((B, self.N))
        elif sforce == 'sin':
            force = sin1
        elif sforce == 'cos':
            force = cos1
        elif sforce == 'sincos':
            pick = self._rng.random(B)[:, None] > 0.5
            force = np.where(pick, sin1, cos1)
        elif sforce == 'fourier':
            r = self._rng.random(B)[:, None]
            force = np.where(r > 0.66, sin1, np.where(r > 0.33, np.sin((x - off) * 3 * np.pi / L), np.sin((x - off) * 4 * np.pi / L)))
        elif sforce == 'gaussian':
            force = np.exp(-0.5 * (0.5 * L - x + off) ** 2)
        else:
            raise SystemExit(f"[Laplace] Error: force {sforce} unknown")
        force = np.broadcast_to(force, (B, self.N)).copy()
        self._u0 = u0
        u0d = self._batch(u0, self.dtype, (self.N,))
        L_check(self._lib.mpde_reset_u(self._h, self._ptr(u0d), None, self._stream()))
        self.u0 = self._squeeze(u0d)
        self._force = force
        shared = B == 1 or bool(np.all(force == force[0]))
        tab = self._dev(force[:1] if shared else force).reshape(-1, 1, self.N).contiguous()
        mp = None
        if not shared:
            self._keep['truth_map'] = self._dev(np.arange(B), torch.int32, (B,))
            mp = self._ptr(self._keep['truth_map'])
        self._keep['truth'] = tab
        L_check(self._lib.mpde_set_truth(self._h, self._ptr(tab), tab.shape[0], 1, mp))
        L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_DIRECT))
        self.t, self.stepnum, self.ioutnum = 0., 0, 0

    @property
    def force(self):
        return self._force[0] if self.nenvs == 1 else self._force

    def step_n(self, actions, numAgents, n=1, want_state=False, want_reward=False):
        assert numAgents + 1 == self.N                                    # Laplace.py:118 (one boundary point)
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=self.dtype)
        else:
            a = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self.device).to(self.dtype)
        a = a.reshape(self.nenvs, -1).contiguous()
        assert a.shape[1] == 3 * numAgents, f"[Laplace] action len not 3, it is {a.shape[1] / numAgents}"
        nA = numAgents
        st = torch.empty((self.nenvs, 4 * nA), device=self.device, dtype=self.dtype) if want_state else None
        rw = torch.empty((self.nenvs, nA), device=self.device, dtype=self.dtype) if want_reward else None
        L_check(self._lib.mpde_step(self._h, self._ptr(a), int(n), self._ptr(st), self._ptr(rw), self._stream()))
        self.stepnum += n
        self.ioutnum += n
        for _ in range(n):
            self.t += self.dt
        return st, rw

    def step(self, actions, numAgents):
        self.step_n(actions, numAgents, 1)

    def getDirectReward(self, numAgents=1, as_tensor=None):
        """Laplace.py:153-160."""
        assert numAgents + 1 == self.N, f"[Laplace] direct reward neeeds N agents (using {numAgents})"
        rw = torch.empty((self.nenvs, numAgents), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, None, 0, None, self._ptr(rw), self._stream()))
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return rw if as_tensor else rw[0].cpu().numpy().tolist()

    def getState(self, numAgents=1, as_tensor=None):
        """Laplace.py:162-166."""
        assert self.N == numAgents + 1
        st = torch.empty((self.nenvs, 4 * numAgents), device=self.device, dtype=self.dtype)
        L_check(self._lib.mpde_step(self._h, None, 0, self._ptr(st), None, self._stream()))
        st = st.view(self.nenvs, numAgents, 4)
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return st if as_tensor else st[0].cpu().numpy().tolist()
