"""Host-side logic shared by the two pseudo-spectral environments (Burger, KS): history
buffers, u / v attributes, the action basis, the spectrum reference of the spectral reward
and the energy-spectrum views.  See Burger.py / KS.py for the reference lines."""
import numpy as np
import torch

from . import _lib as LB
from ._base import BatchedEnv
from .hostmath import make_basis

L_check = LB.check


class SpectralEnv(BatchedEnv):
    def _squeeze(self, t):
        return t[0] if self.nenvs == 1 else t

    @property
    def u(self):
        return self._squeeze(self._get(LB.FIELD_U, (self.nenvs, self.N), self.dtype))

    @property
    def v(self):
        return self._squeeze(self._get(LB.FIELD_V, (self.nenvs, self.N), self.cdtype))


    # ------------------------------------------------------------------ device-side episode reset
    def IC_handoff(self, src_v0, src_k, src_map=None, offsets=None, mask=None):
        """DNS -> LES spectral hand-off for the whole batch in one launch (burger_environment.py:109-112,
        ks_environment.py:52-54):  v0_e = concat(w[:(N+1)//2], w[-(N-1)//2:]) * N / N_src with
        w = src_v0[src_map[e]] * exp(1j * 2 pi * offsets[e] * src_k), then IC(v0=v0_e).
        src_v0: [N_src] or [n_src, N_src] complex (dns.v0 of one or several DNS runs), src_k: [N_src] (dns.k)."""
        B = self.nenvs
        v = torch.as_tensor(src_v0) if not isinstance(src_v0, torch.Tensor) else src_v0
        v = v.to(device=self.device, dtype=torch.complex128).reshape(-1, v.shape[-1]).contiguous()
        k = torch.as_tensor(np.asarray(src_k, dtype=np.float64), device=self.device).contiguous()
        assert k.numel() == v.shape[1]
        mp_t = None if src_map is None else self._dev(np.asarray(src_map), torch.int32, (B,))
        off_t = None if offsets is None else self._dev(np.broadcast_to(np.asarray(offsets, dtype=np.float64), (B,)).copy(),
                                                       torch.float64, (B,))
        m, mp = self._mask_ptr(mask)
        L_check(self._lib.mpde_reset_handoff(self._h, self._ptr(torch.view_as_real(v)), v.shape[0], v.shape[1], self._ptr(k),
                                             self._ptr(mp_t), self._ptr(off_t), mp, self._stream()))
        self._after_reset(mask)

    def _after_reset(self, mask=None):
        """Host-side bookkeeping after an IC.  The device keeps ``ioutnum`` / ``t`` PER ENVIRONMENT (the kernels use those
        for the forcing column and the reward row); the host scalars ``t``, ``ioutnum``, ``stepnum`` describe the whole
        batch and are only meaningful while every environment is at the same step.  A MASKED reset of a running batch
        breaks that: the scalars are left alone, ``_out_of_step`` is set, and everything that slices by the host scalar
        (``compute_Ek``, ``uu`` / ``vv`` history views, ``state_dict``) raises until a full reset; per-environment
        counters stay available through ``ioutnum_all`` / ``status``."""
        partial = mask is not None and not bool(np.all(np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask) != 0))
        if partial and getattr(self, "ioutnum", 0) != 0:
            self._out_of_step = True
            self._state_at = self._reward_at = -1
            if hasattr(self, "_uu_valid_at"):
                self._uu_valid_at = -1
            return
        self._out_of_step = False
        self.t = 0.
        self.stepnum = 0
        self.ioutnum = 0
        self._state_at = self._reward_at = -1
        if hasattr(self, "_uu_valid_at"):
            self._uu_valid_at = -1
        self.u0 = self.u
        self.v0 = self.v

    # ------------------------------------------------------------------ history
    def _setup_history(self, history):
        B, rows, N = self.nenvs, self.nout + 1, self.N
        per_row = N * (torch.empty((), dtype=self.dtype).element_size() + 8) + (N // 2 + 1) * 8
        if history is None:
            history = B * rows * per_row <= (2 << 30)
        self.history = bool(history)
        self.tt = np.concatenate(([0.], np.cumsum(np.full(self.nout, self.dt))))   # t += dt (Burger.py:494,499)
        if self.history:
            self._uu = torch.zeros((B, rows, N), device=self.device, dtype=self.dtype)
            self._vv = torch.zeros((B, rows, N), device=self.device, dtype=torch.complex64)
            self._ektt = torch.zeros((B, rows, N // 2 + 1), device=self.device, dtype=torch.float64)
            L_check(self._lib.mpde_set_history(self._h, self._ptr(self._uu), self._ptr(self._vv),
                                               self._ptr(self._ektt), rows))
        else:
            self._uu = self._vv = self._ektt = None
            L_check(self._lib.mpde_set_history(self._h, None, None, None, 0))

    @property
    def uu(self):
        self._need_history()
        return self._squeeze(self._uu)

    @property
    def vv(self):
        self._need_history()
        return self._squeeze(self._vv)

    def compute_Sgs(self, nURG):
        """Burger.py:677-736 / KS.py:385-409: a-priori sub-grid-scale term of the recorded history for a coarse grid of
        ``nURG`` modes (testing-mode diagnostic).  Sets ``sgsHistory`` [rows, N] and, for Burgers, ``sgsHistoryAlt``
        [rows, N], ``sgsHistoryAlt2`` [rows, nURG] (device tensors; leading environment axis when nenvs > 1)."""
        self._need_history()
        B, rows, N = self.nenvs, self.nout + 1, self.N
        burgers = self.equation == LB.BURGERS
        sgs = torch.zeros((B, rows, N), device=self.device, dtype=self.dtype)
        alt = torch.zeros((B, rows, N), device=self.device, dtype=self.dtype) if burgers else None
        alt2 = torch.zeros((B, rows, int(nURG)), device=self.device, dtype=self.dtype) if burgers else None
        L_check(self._lib.mpde_compute_sgs(self._h, int(nURG), rows, self._ptr(sgs), self._ptr(alt), self._ptr(alt2), self._stream()))
        self.sgsHistory = self._squeeze(sgs)
        if burgers:
            self.sgsHistoryAlt, self.sgsHistoryAlt2 = self._squeeze(alt), self._squeeze(alt2)

    def _need_in_step(self, what):
        if getattr(self, "_out_of_step", False):
            raise RuntimeError(f"{what}: a masked reset left the environments of this batch at different steps; the host-side "
                               "scalars (t, ioutnum) no longer describe them -- use ioutnum_all / status, or reset the whole batch")

    def _need_history(self):
        self._need_in_step("history / spectrum views")
        if not self.history:
            raise RuntimeError("history recording is off for this batch (pass history=True)")

    def setup_basis(self, M, kind='uniform'):
        """Burger.py:177-203 / KS.py:139-164."""
        self.M = M
        self.basis = make_basis(self.x, self.L, M, kind)
        L_check(self._lib.mpde_set_basis(self._h, int(M), LB.as_dp(np.ascontiguousarray(self.basis))))

    def set_spectrum_reference(self, ref, env_map=None):
        """Reference spectrum rows for the spectral reward (burger_environment.py:174):
        a DNS ``Burger`` with history, or an array [rows, >=N/2] / [nref, rows, >=N/2]."""
        h = self.N // 2
        if isinstance(ref, SpectralEnv):
            ref._need_history()
            tab = ref._ektt[:, :, :h]
        else:
            tab = self._dev(ref, torch.float64)
            if tab.dim() == 2:
                tab = tab.unsqueeze(0)
            tab = tab[:, :, :h]
        tab = tab.to(self.device).contiguous()
        mp = None
        if env_map is not None:
            self._keep['ek_map'] = self._dev(env_map, torch.int32, (self.nenvs,))
            mp = self._ptr(self._keep['ek_map'])
        self._keep['ek_ref'] = tab
        L_check(self._lib.mpde_set_spectrum_ref(self._h, self._ptr(tab), tab.shape[0], tab.shape[1], mp))
        L_check(self._lib.mpde_set_reward_mode(self._h, LB.REWARD_SPECTRAL))
        self._spec_ref = tab

    def Ek_ktt_row(self):
        """Row ``ioutnum`` of Ek_ktt[:, :N/2+1] (Burger.py:555) from the running float32 sums."""
        acc = self._get(LB.FIELD_EK_SUM, (self.nenvs, self.N // 2 + 1), torch.float32)
        cnt = (self.ioutnum_all + 1).to(torch.float64).unsqueeze(1)
        return self._squeeze(acc.to(torch.float64) / cnt)

    def compute_Ek(self):
        """Burger.py:541-576 from the recorded history (float32 chain of the complex64 vv)."""
        self._need_history()
        i = self.ioutnum
        vv = self._vv[:, :i + 1]
        self.Ek_kt = self._squeeze(0.5 * torch.real(vv.conj() * vv / self.N) * np.float32(self.dx))
        ekt = self.Ek_kt if self.nenvs > 1 else self.Ek_kt[None]
        self.Ek_k = self._squeeze(ekt.sum(1) / (i + 1))
        self.Ek_t = self._squeeze(ekt.sum(2))
        n = self.N
        half = self._ektt[:, :i + 1]                                       # exact sequential float32 sums
        idx = torch.arange(n, device=self.device)
        idx = torch.where(idx <= n // 2, idx, n - idx)
        self.Ek_ktt = self._squeeze(half[:, :, idx])
        den = torch.arange(1, i + 2, device=self.device, dtype=torch.float64)
        self.Ek_tt = self._squeeze(torch.cumsum(ekt.sum(2), 1).to(torch.float64) / den)

