"""Host-side plumbing shared by the batched environment classes.

PyTorch is used for device memory, streams and (optionally) torch.distributed only; all
solver arithmetic runs in libmarlpde_b200.so.  A missing library or a missing GPU raises
-- there is no eager/CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _torch_dtype(dtype):
    if dtype in (torch.float64, "float64", "f64", np.float64, float):
        return torch.float64
    if dtype in (torch.float32, "float32", "f32", np.float32):
        return torch.float32
    raise ValueError(f"unsupported dtype {dtype!r} (float64 or float32)")


class BatchedEnv:
    """Owns one ``mpde_env`` handle bound to one CUDA device."""

    equation = None

    def _create(self, *, nenvs, N, L_, dt, M=0, num_agents=1, version=0, stepper=1, flags=0,
                reward_mode=L.REWARD_NONE, device=None, dtype=torch.float64, team_lanes=0):
        if not torch.cuda.is_available():
            raise RuntimeError("marlpde_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self._lib = L.lib()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("marlpde_b200 environments live on a CUDA device")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.dtype = _torch_dtype(dtype)
        self.cdtype = torch.complex128 if self.dtype == torch.float64 else torch.complex64
        self.nenvs = int(nenvs)
        cfg = L.MpdeConfig()
        cfg.struct_size = C.sizeof(L.MpdeConfig)
        cfg.equation = self.equation
        cfg.dtype = L.F64 if self.dtype == torch.float64 else L.F32
        cfg.device = device.index
        cfg.nenvs = self.nenvs
        cfg.N, cfg.M, cfg.num_agents, cfg.version = int(N), int(M), int(num_agents), int(version)
        cfg.stepper, cfg.flags, cfg.reward_mode = int(stepper), int(flags), int(reward_mode)
        cfg.L, cfg.dt = float(L_), float(dt)
        cfg.team_lanes = int(team_lanes or 0)
        h = C.c_void_p()
        L.check(self._lib.mpde_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._keep = {}          # device tensors the library holds raw pointers to
        self._state_size = int(self._lib.mpde_state_size(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.mpde_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- helpers ------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x, dtype=None, shape=None):
        """Anything array-like -> contiguous device tensor of the env dtype."""
        dtype = dtype or self.dtype
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=dtype)
        else:
            t = torch.as_tensor(np.asarray(x), device=self.device).to(dtype)
        if shape is not None:
            t = t.reshape(shape)
        return t.contiguous()

    def _ptr(self, t):
        if t is None:
            return None
        assert t.is_cuda and t.is_contiguous() and t.device == self.device, "need a contiguous tensor on the env device"
        return C.c_void_p(t.data_ptr())

    def _batch(self, x, dtype, tail):
        """Broadcast a single-env array to [B, *tail] on the device."""
        t = self._dev(x, dtype)
        if t.dim() == len(tail):
            t = t.unsqueeze(0).expand(self.nenvs, *t.shape)
        return t.reshape(self.nenvs, *tail).contiguous()

    def _get(self, field, shape, dtype):
        out = torch.empty(shape, device=self.device, dtype=dtype)
        L.check(self._lib.mpde_get(self._h, field, self._ptr(out), self._stream()))
        return out

    def _set(self, field, t):
        L.check(self._lib.mpde_set(self._h, field, self._ptr(t), self._stream()))

    def _mask_ptr(self, mask):
        if mask is None:
            return None, None
        m = self._dev(mask, torch.uint8, (self.nenvs,))
        return m, self._ptr(m)

    # ---- attributes common to every solver -------------------------------------------
    @property
    def status(self):
        """[B] int32: 0 running, 1 truncated (numerical blow-up; the reference raises
        FloatingPointError and the environment reports "Truncated")."""
        return self._get(L.FIELD_STATUS, (self.nenvs,), torch.int32)

    @property
    def ioutnum_all(self):
        return self._get(L.FIELD_IOUTNUM, (self.nenvs,), torch.int32)

    @property
    def launch_count(self):
        return int(self._lib.mpde_launch_count(self._h))

    def _scalar_or_tensor(self, t):
        return t[0].item() if self.nenvs == 1 else t
