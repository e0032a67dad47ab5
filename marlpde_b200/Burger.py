"""Batched, GPU-resident drop-in for the reference ``Burger`` environment class.

Mirrors the public API of /root/reference/python/_model/Burger.py (constructor keywords
:24-44, ``setup_basis`` :177, ``IC`` :205, ``setGroundTruth`` :322, ``step`` :333,
``simulate`` :501, ``compute_Ek`` :541, ``getMseReward`` :578, ``getState`` :604) for
``nenvs`` independent environments that live on one B200.  With ``nenvs=1`` the methods
return exactly what the reference returns (nested lists / numpy); with ``nenvs>1`` they
return device tensors with the environment index first.

Only host-side set-up (initial conditions, forcing tables, basis, truth interpolation) is
computed here with numpy; every ``step`` runs in the CUDA library (marlpde_b200/csrc).
"""
import numpy as np
import torch

from . import _lib as LB
from ._spectral import SpectralEnv
from .hostmath import (grid_points, fft_wavenumbers, turbulence_field, forced_field,
                       forcing_spectrum_coefficients, TruthInterpolant)


class Burger(SpectralEnv):
    equation = LB.BURGERS

    def __init__(self, L=2. * np.pi, N=1024, dt=0.001, nu=0.02, dforce=True, ssmforce=False, nsteps=None,
                 tend=5., u0=None, v0=None, case=None, forcing=False, ssm=False, dsm=False, noise=0., seed=42,
                 version=0, nunoise=False, numAgents=1, s=1, *, nenvs=1, device=None, dtype=torch.float64,
                 history=None, offset=None, team_lanes=0):
        assert not (ssm and dsm)                                            # Burger.py:50
        if ssmforce and not dforce:
            raise SystemExit("[Burger] SSM forcing requires dforce")      # Burger.py:113-115
        B = int(nenvs)
        self.numAgents = numAgents
        self.L, self.dt, self.tend = float(L), float(dt), float(tend)
        nsteps = int(tend / dt) if nsteps is None else int(nsteps)
        self.N, self.dx = N, L / N
        self.x = grid_points(self.L, N)
        self.nsteps = self.nout = nsteps
        self.stepper = int(s)
        self.forcing, self.ssm, self.dsm = bool(forcing), bool(ssm), bool(dsm)
        self.dforce, self.ssmforce = bool(dforce), bool(ssmforce)
        self.version, self.cs = version, 0.1
        self.M, self.basis, self.f_truth = 0, None, None

        # per-environment seeds; offset ~ N(0, noise*L) redrawn until |offset| <= L (Burger.py:53-57;
        # the reference draws it from an UNSEEDED generator -- pass offset= to pin it)
        self.seeds = np.broadcast_to(np.asarray(seed, dtype=np.int64), (B,)).copy()
        self.tseed = self.seeds
        self.noise = noise * L
        if offset is not None:
            off = np.broadcast_to(np.asarray(offset, dtype=np.float64), (B,)).copy()
        elif self.noise > 0.:
            rng = np.random.default_rng()
            off = rng.normal(0., self.noise, B)
            while np.any(np.abs(off) > L):
                bad = np.abs(off) > L
                off[bad] = rng.normal(0., self.noise, int(bad.sum()))
        else:
            off = np.zeros(B)
        self._offset = off

        # seeded stream: [nunoise uniform] -> randfac1 -> randfac2 -> ('forced' IC draws) (Burger.py:66,88-95)
        nus = np.full(B, float(nu))
        self._streams = {}
        need_tables = self.forcing or case == 'forced' or nunoise
        uniq = np.unique(self.seeds)
        r1 = r2 = None
        self._tables_on_device = False
        if need_tables and len(uniq) > 16 and case != 'forced' and int(uniq.min()) >= 0 and int(uniq.max()) < 2 ** 32:
            # many distinct seeds (seed = 42 + e): the NumPy stream of every seed is reproduced on the device
            # (mpde_forcing_tables, one warp per seed); only rows 1..3 / columns < stepper exist -- all the solver reads
            # (Burger.py:416-419) -- the other rows of the [B, 32, stepper] tables are NaN
            nu_u, a3, b3 = device_forcing_tables(uniq, nsteps, self.stepper, nunoise, device)
            inv = np.searchsorted(uniq, self.seeds)
            r1 = np.full((B, 32, self.stepper), np.nan)
            r2 = np.full((B, 32, self.stepper), np.nan)
            r1[:, 1:4], r2[:, 1:4] = a3[inv], b3[inv]
            if nunoise:
                nus = nu_u[inv]
            self._tables_on_device = True
        elif need_tables:
            per_seed = {}
            for sd in uniq:
                rs = np.random.RandomState(int(sd))
                nu_s = 0.01 + 0.02 * rs.uniform() if nunoise else float(nu)
                a = rs.normal(loc=0., scale=1., size=(32, nsteps))
                b = rs.normal(loc=0., scale=1., size=(32, nsteps))
                per_seed[int(sd)] = (nu_s, a, b)
                self._streams[int(sd)] = rs
            if nunoise:
                nus = np.array([per_seed[int(sd)][0] for sd in self.seeds])
            if len(uniq) == 1:
                r1, r2 = per_seed[int(uniq[0])][1:]
            else:   # keep only the columns the stepper can ever index (ioutnum % s, Burger.py:416)
                r1 = np.stack([per_seed[int(sd)][1][:, :self.stepper] for sd in self.seeds])
                r2 = np.stack([per_seed[int(sd)][2][:, :self.stepper] for sd in self.seeds])
        self._randfac1, self._randfac2 = r1, r2
        self._nu = nus
        self._forcing_dirty = True

        flags = (L_DFORCE if dforce else 0) | (L_FORCING if forcing else 0) | (L_SSM if ssm else 0) | (L_DSM if dsm else 0)
        flags |= self._extra_flags()
        self._create(nenvs=B, N=N, L_=self.L, dt=self.dt, M=0, num_agents=numAgents, version=version,
                     stepper=self.stepper, flags=flags, device=device, dtype=dtype, team_lanes=team_lanes)
        L_check(self._lib.mpde_set_nu(self._h, LB.as_dp(np.ascontiguousarray(self._nu)), B))

        self.k = fft_wavenumbers(self.L, N)                                # Burger.py:161-163
        self.k1 = 1j * self.k
        self.k2 = self.k1 ** 2

        self._setup_history(history)
        self._state_buf = torch.empty((B, self._state_size), device=self.device, dtype=self.dtype)
        self._reward_buf = torch.zeros((B, numAgents), device=self.device, dtype=self.dtype)
        self._state_at = self._reward_at = -1
        self._spec_ref = None
        self._truth_shift = None

        if case is not None:
            self.IC(case=case)
        elif u0 is None and v0 is None:
            self.IC()
        elif u0 is not None:
            self.IC(u0=u0)
        else:
            self.IC(v0=v0)

    def _extra_flags(self):
        return 0

    # ------------------------------------------------------------------ simple attributes
    @property
    def nu(self):
        return float(self._nu[0]) if self.nenvs == 1 else self._nu

    @property
    def offset(self):
        return float(self._offset[0]) if self.nenvs == 1 else self._offset

    @offset.setter
    def offset(self, value):
        self._offset = np.broadcast_to(np.asarray(value, dtype=np.float64), (self.nenvs,)).copy()
        self._forcing_dirty = True

    @property
    def randfac1(self):
        return self._randfac1

    @randfac1.setter
    def randfac1(self, value):           # the environment copies the DNS tables (burger_environment.py:99-100)
        self._randfac1 = np.asarray(value)
        self._forcing_dirty = True

    @property
    def randfac2(self):
        return self._randfac2

    @randfac2.setter
    def randfac2(self, value):
        self._randfac2 = np.asarray(value)
        self._forcing_dirty = True

    @property
    def Fn_old(self):
        return self._squeeze(self._get(LB.FIELD_FN_OLD, (self.nenvs, self.N), self.cdtype))

    # ------------------------------------------------------------------ set-up
    def IC(self, u0=None, v0=None, case='zero', mask=None, on_device=None):
        """Burger.py:205-320.  u0 / v0: [N] (shared) or [B, N]; ``mask`` resets a subset.
        case='turbulence' with many distinct (seed, offset) pairs is generated on the device (one CTA per
        environment, mpde_reset_turbulence) instead of in a host loop; ``on_device`` forces either path."""
        B, N = self.nenvs, self.N
        if v0 is None and u0 is None and case == 'turbulence':
            if on_device is None:
                on_device = len({(int(a), float(b)) for a, b in zip(self.seeds, self._offset)}) > 16
            if on_device:
                k = np.arange(N, dtype=np.float64)
                with np.errstate(divide='ignore'):
                    Ek = np.where(k <= 5, 5. ** (-5 / 3), k ** (-5 / 3))            # Burger.py:242
                amp = np.sqrt(2 * Ek)
                sd = self._dev(self.seeds, torch.int64, (B,))
                off = self._dev(self._offset, torch.float64, (B,))
                xd = self._dev(self.x, torch.float64, (N,))
                ad = self._dev(amp, torch.float64, (N,))
                m, mp = self._mask_ptr(mask)
                L_check(self._lib.mpde_reset_turbulence(self._h, self._ptr(sd), self._ptr(off), self._ptr(xd), self._ptr(ad), mp,
                                                        self._stream()))
                self._after_reset(mask)
                return
        if v0 is None:
            if u0 is None:
                u0 = self._case_field(case)
            else:
                if np.shape(u0)[-1] != N:
                    raise SystemExit(f"[Burger] Error: wrong IC array size (is {np.shape(u0)[-1]}, expected {N}")
            u0d = self._batch(u0, self.dtype, (N,))
            m, mp = self._mask_ptr(mask)
            L_check(self._lib.mpde_reset_u(self._h, self._ptr(u0d), mp, self._stream()))
        else:
            if np.shape(v0)[-1] != N:
                raise SystemExit(f"[Burger] Error: wrong IC array size (is {np.shape(v0)[-1]}, expected {N}")
            v0d = self._batch(v0, self.cdtype, (N,))
            m, mp = self._mask_ptr(mask)
            L_check(self._lib.mpde_reset_v(self._h, self._ptr(torch.view_as_real(v0d)), mp, self._stream()))
        self._after_reset(mask)

    def _case_field(self, case):
        B, N = self.nenvs, self.N
        out = np.empty((B, N))
        cache = {}
        for e in range(B):
            key = (int(self.seeds[e]), float(self._offset[e]))
            if key not in cache:
                if case == 'sinus':                                        # Burger.py:224
                    f = np.sin(4. * np.pi * (self.x + key[1]) / self.L)
                elif case == 'turbulence':                                 # Burger.py:227-260
                    f = turbulence_field(self.x, self.L, N, key[1], key[0])
                elif case == 'zero':
                    f = np.zeros(N)
                elif case == 'forced':                                     # Burger.py:265-273
                    f = forced_field(self.x, self.L, N, self._streams[key[0]])
                else:
                    raise SystemExit("[Burger] Error: IC case unknown")
                cache[key] = f
            out[e] = cache[key]
        return out

    def setGroundTruth(self, x, t, uu_truth):
        """Burger.py:322-323 (argument order x, t, uu).  The cubic tensor-product spline is
        built lazily on the host; the device only sees the truth sampled on this grid."""
        self.f_truth = TruthInterpolant(_np(x), _np(t), _np(uu_truth), kind='cubic')
        self._truth_shift = None

    def mapGroundTruth(self):
        t = np.arange(0, self.nout + 1) * self.dt
        return self.f_truth(self.x, t)

    def set_truth_table(self, table, env_map=None):
        """Directly provide the truth sampled at this grid: [rows, N] or [ntruth, rows, N]."""
        tab = self._dev(table)
        if tab.dim() == 2:
            tab = tab.unsqueeze(0)
        mp = None
        if env_map is not None:
            self._keep['truth_map'] = self._dev(env_map, torch.int32, (self.nenvs,))
            mp = self._ptr(self._keep['truth_map'])
        self._keep['truth'] = tab.contiguous()
        L_check(self._lib.mpde_set_truth(self._h, self._ptr(self._keep['truth']), tab.shape[0], tab.shape[1], mp))
        L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_MSE))
        self._truth_shift = 'explicit'

    def _ensure_truth(self, shift):
        """Sample the spline at x + shift (wrapped into [0, L], Burger.py:581-588) for every
        row of tt; environments sharing a shift share a table."""
        shift = np.broadcast_to(np.asarray(shift, dtype=np.float64), (self.nenvs,))
        key = shift.tobytes()
        if self._truth_shift == key or self._truth_shift == 'explicit':
            return
        if self.f_truth is None:
            raise RuntimeError("getMseReward needs setGroundTruth(...) or set_truth_table(...)")
        uniq, inv = np.unique(shift, return_inverse=True)
        grids = []
        for sh in uniq:
            newx = self.x + sh
            newx[newx > self.L] -= self.L
            newx[newx < 0] += self.L
            grids.append(newx)
        # fit and sample the spline on the device: no host FITPACK call on the reward path
        tabs = self.f_truth.rows_device(np.stack(grids), self.tt, self.device, self.dtype)
        self.set_truth_table(tabs, env_map=inv.astype(np.int32) if len(uniq) > 1 else None)
        self._truth_shift = key

    def _upload_forcing(self):
        if not self.forcing:
            self._forcing_dirty = False
            return
        if not self._forcing_dirty:
            return
        coef = forcing_spectrum_coefficients(self._randfac1, self._randfac2, self._offset, self.L, self.dt,
                                             self.stepper, self.N, self.nenvs)
        L_check(self._lib.mpde_set_forcing(self._h, LB.as_dp(coef), coef.shape[0]))
        self._forcing_dirty = False

    # ------------------------------------------------------------------ stepping
    def _actions(self, actions):
        if actions is None:
            return None
        assert self.basis is not None, "[Burger] Basis not set up (is None)."
        if (isinstance(actions, torch.Tensor) and actions.device == self.device and actions.dtype == self.dtype and actions.dim() == 2
                and actions.shape[0] == self.nenvs and actions.shape[1] == self.M and actions.is_contiguous()):
            return actions                                                 # the learner's own [B,M] device tensor: no copy
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=self.dtype)
        else:
            a = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self.device).to(self.dtype)
        a = a.reshape(self.nenvs, -1)                                      # MARL lists are flattened (Burger.py:437)
        assert a.shape[1] == self.M, "[Burger] Wrong number of actions (provided {}/{}".format(a.shape[1], self.M)
        return a.contiguous()

    def bind_output(self, state_buf, reward_buf):
        """Let the caller own the [B,S] state and [B,A] reward buffers step_n writes (e.g. two views of
        one flat tensor that is all-gathered to the learner in a single collective)."""
        assert state_buf.shape == self._state_buf.shape and reward_buf.shape == self._reward_buf.shape
        assert state_buf.is_contiguous() and reward_buf.is_contiguous()
        self._state_buf, self._reward_buf = state_buf, reward_buf
        self._state_at = self._reward_at = -1

    def step_n(self, actions=None, n=1, want_state=True, want_reward=True):
        """``n`` solver steps with the same actions (the inner loop of
        burger_environment.py:148-155) + getState + reward, as ONE kernel launch.
        Returns (state [B,S], reward [B,A]) device tensors (None when not requested)."""
        if self._forcing_dirty:
            self._upload_forcing()
        a = self._actions(actions)
        st = self._state_buf if want_state else None
        rw = self._reward_buf if (want_reward and (self._spec_ref is not None or self._truth_shift is not None)) else None
        # raw addresses (ctypes converts ints to void*): this call sits on the host path of every RL step
        rc = self._lib.mpde_step(self._h, a.data_ptr() if a is not None else None, int(n),
                                 st.data_ptr() if st is not None else None, rw.data_ptr() if rw is not None else None,
                                 torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            L_check(rc)
        self.stepnum += n
        self.ioutnum += n
        t, dt = self.t, self.dt
        for _ in range(n):
            t += dt                                                        # Burger.py:494 (accumulated, not n * dt)
        self.t = t
        if st is not None:
            self._state_at = self.ioutnum
        if rw is not None:
            self._reward_at = self.ioutnum
        return st, rw

    def _reward_enabled(self):
        return self._spec_ref is not None or self._truth_shift is not None

    def step_n_fused(self, actions, n, async_gather=False):
        """Multi-GPU form of step_n with a fused gather bound (dist.PeerGather.fuse): step kernel storing the rows into every
        rank's buffer -> publish -> wait for all ranks, as ONE library call replayed from a cached CUDA graph
        (mpde_step_fused).  ``async_gather``: the publish / wait pair runs on a side stream; ``peer_join()`` orders the
        current stream behind it."""
        self._upload_forcing()
        a = self._actions(actions)
        L_check(self._lib.mpde_step_fused(self._h, self._ptr(a), int(n), self._ptr(self._state_buf), self._ptr(self._reward_buf),
                                          1 if async_gather else 0, self._stream()))
        self.stepnum += n
        self.ioutnum += n
        for _ in range(n):
            self.t += self.dt
        self._state_at = self._reward_at = -1                              # this step's rows live in the gather copy
        return None

    def peer_join(self):
        L_check(self._lib.mpde_peer_join(self._h, self._stream()))

    def step_n_host(self, actions_host, n, state_host, reward_host=None, stream=None, packed_out=None):
        """Host-buffer form of step_n for a host-side policy: (pinned) host actions [B,M] in, (pinned) host
        state [B,S] / reward [B,A] out, everything enqueued asynchronously on ``stream`` (default: the current
        stream) by ONE library call (H2D copy -> step kernel -> D2H copies, replayed from a cached CUDA graph).
        ``packed_out``: one pinned buffer [B*S | B*A] instead of state_host / reward_host -- a single D2H copy.
        Results are valid after the stream / an event recorded behind this call completes."""
        self._upload_forcing()
        for t_ in (actions_host, state_host, reward_host, packed_out):
            assert t_ is None or (not t_.is_cuda and t_.is_contiguous() and t_.dtype == self.dtype)
        st = stream.cuda_stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        have_rw = self._spec_ref is not None or self._truth_shift is not None
        a_ptr = actions_host.data_ptr() if actions_host is not None else None
        if packed_out is not None:
            assert have_rw and packed_out.numel() == self.nenvs * (self._state_size + self._reward_buf.shape[1])
            L_check(self._lib.mpde_step_host_packed(self._h, a_ptr, int(n), packed_out.data_ptr(), st))
        else:
            L_check(self._lib.mpde_step_host(self._h, a_ptr, int(n), state_host.data_ptr() if state_host is not None else None,
                                             reward_host.data_ptr() if (have_rw and reward_host is not None) else None, st))
        self.stepnum += n
        self.ioutnum += n
        for _ in range(n):
            self.t += self.dt
        self._state_at = self._reward_at = -1

    def host_step_closure(self, actions_host, n, packed_out, stream):
        """``step_n_host(actions_host, n, ..., packed_out=packed_out, stream=stream)`` with everything that does not change
        between calls resolved once: returns a zero-argument callable for a learner's inner loop (one ctypes call into
        ``mpde_step_host_packed`` + the host counters; about a third of the Python time of ``step_n_host``)."""
        for t_ in (actions_host, packed_out):
            assert not t_.is_cuda and t_.is_contiguous() and t_.dtype == self.dtype
        assert self._spec_ref is not None or self._truth_shift is not None
        assert packed_out.numel() == self.nenvs * (self._state_size + self._reward_buf.shape[1])
        fn, h, n = self._lib.mpde_step_host_packed, self._h, int(n)
        a_ptr, o_ptr, st, dt = actions_host.data_ptr(), packed_out.data_ptr(), stream.cuda_stream, self.dt

        def go():
            if self._forcing_dirty:
                self._upload_forcing()
            if fn(h, a_ptr, n, o_ptr, st) != 0:
                L_check(-1)
            self.stepnum += n
            self.ioutnum += n
            for _ in range(n):
                self.t += dt
            self._state_at = self._reward_at = -1
        return go

    def step(self, actions=None):
        """Burger.py:333-499: one solver step."""
        self.step_n(actions, 1, want_state=False, want_reward=self._truth_shift is not None)

    def simulate(self, nsteps=None, restart=False, correction=[]):
        """Burger.py:501-539 (``correction`` is not supported)."""
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
            n = min(left, 500)
            self.step_n(None, n, want_state=False, want_reward=False)
            left -= n
        if bool((self.status != 0).any()):
            print("[Burger] Floating point exception occured in simulate", flush=True)
            return -1

    # ------------------------------------------------------------------ observables
    def getState(self, nAgents=None, as_tensor=None):
        """Burger.py:604-675.  nenvs == 1: the reference's nested lists; else [B,S] / [B,A,S/A]."""
        if self._state_at != self.ioutnum:
            self._upload_forcing()
            L_check(self._lib.mpde_step(self._h, None, 0, self._ptr(self._state_buf), None, self._stream()))
            self._state_at = self.ioutnum
        A = self.numAgents
        st = self._state_buf if A == 1 else self._state_buf.view(self.nenvs, A, -1)
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        if as_tensor:
            return st
        host = st[0].cpu().numpy()
        return [host.tolist()] if A == 1 else [row.tolist() for row in host]

    def getMseReward(self, shift=0., as_tensor=None):
        """Burger.py:578-601 -> rewards per agent; nenvs == 1: numpy [A]."""
        self._ensure_truth(shift)
        if self._reward_at != self.ioutnum:
            self._upload_forcing()
            L_check(self._lib.mpde_step(self._h, None, 0, None, self._ptr(self._reward_buf), self._stream()))
            self._reward_at = self.ioutnum
        if as_tensor is None:
            as_tensor = self.nenvs > 1
        return self._reward_buf if as_tensor else self._reward_buf[0].cpu().numpy()

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self):
        self._need_in_step("state_dict")
        B, N = self.nenvs, self.N
        return dict(v=self._get(LB.FIELD_V, (B, N), self.cdtype), Fn_old=self._get(LB.FIELD_FN_OLD, (B, N), self.cdtype),
                    u_prev=self._get(LB.FIELD_U_PREV, (B, N), self.dtype),
                    ek_sum=self._get(LB.FIELD_EK_SUM, (B, N // 2 + 1), torch.float32),
                    ioutnum=self._get(LB.FIELD_IOUTNUM, (B,), torch.int32), t=self._get(LB.FIELD_T, (B,), self.dtype),
                    kprev=self._get(LB.FIELD_KPREV, (B,), self.dtype), status=self._get(LB.FIELD_STATUS, (B,), torch.int32),
                    host=dict(t=self.t, ioutnum=self.ioutnum, stepnum=self.stepnum))

    def load_state_dict(self, sd):
        self._set(LB.FIELD_V, torch.view_as_real(sd['v'].to(self.device).contiguous()))
        self._set(LB.FIELD_FN_OLD, torch.view_as_real(sd['Fn_old'].to(self.device).contiguous()))
        self._set(LB.FIELD_U_PREV, sd['u_prev'].to(self.device).contiguous())
        self._set(LB.FIELD_EK_SUM, sd['ek_sum'].to(self.device).contiguous())
        self._set(LB.FIELD_IOUTNUM, sd['ioutnum'].to(self.device).contiguous())
        self._set(LB.FIELD_T, sd['t'].to(self.device).contiguous())
        self._set(LB.FIELD_KPREV, sd['kprev'].to(self.device).contiguous())
        self._set(LB.FIELD_STATUS, sd['status'].to(self.device).contiguous())
        self.t, self.ioutnum, self.stepnum = sd['host']['t'], sd['host']['ioutnum'], sd['host']['stepnum']
        self._state_at = self._reward_at = -1


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def device_forcing_tables(seeds, nsteps, stepper, nunoise=False, device=None):
    """randfac1[1:4, :stepper], randfac2[1:4, :stepper] (and nu when ``nunoise``) of ``np.random.seed(seed)`` for every
    seed, generated on the device (Burger.py:66,88-95; mpde_forcing_tables).  Returns (nu [n] or None, r1 [n,3,s], r2 [n,3,s])."""
    lib = LB.lib()
    if not torch.cuda.is_available():
        raise RuntimeError("marlpde_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    sd = torch.as_tensor(np.asarray(seeds, dtype=np.int64), device=dev).contiguous()
    n = sd.numel()
    r1 = torch.empty((n, 3, stepper), dtype=torch.float64, device=dev)
    r2 = torch.empty_like(r1)
    nu = torch.empty((n,), dtype=torch.float64, device=dev) if nunoise else None
    with torch.cuda.device(dev):
        rc = lib.mpde_forcing_tables(sd.data_ptr(), n, int(nsteps), int(stepper), 1 if nunoise else 0, r1.data_ptr(), r2.data_ptr(),
                                     nu.data_ptr() if nunoise else None, torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise RuntimeError("marlpde_b200: " + lib.mpde_rng_last_error().decode())
    return (nu.cpu().numpy() if nunoise else None), r1.cpu().numpy(), r2.cpu().numpy()


L_DFORCE, L_FORCING, L_SSM, L_DSM = LB.DFORCE, LB.FORCING, LB.SSM, LB.DSM
L_check = LB.check
