"""ctypes binding of libmarlpde_b200.so (C ABI declared in include/marlpde_b200.h).

There is NO fallback: if the shared library is missing or fails to load, every solver
class raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C marlpde_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPDE_LIB_PATH") or os.path.join(_HERE, "libmarlpde_b200.so")     # override: tuning experiments only

# enums (include/marlpde_b200.h)
BURGERS, KS, DIFFUSION, ADVECTION, DIFFUSION_ERROR, LAPLACE = 0, 1, 2, 3, 4, 5
F64, F32 = 0, 1
RUNNING, TRUNCATED = 0, 1
REWARD_NONE, REWARD_SPECTRAL, REWARD_MSE, REWARD_DIRECT = 0, 1, 2, 3
DFORCE, FORCING, SSM, DSM, IMPLICIT, FD, SSMFORCE = 1, 2, 4, 8, 16, 32, 64
(FIELD_U, FIELD_V, FIELD_FN_OLD, FIELD_U_PREV, FIELD_EK_SUM, FIELD_IOUTNUM, FIELD_T, FIELD_KPREV,
 FIELD_STATUS, FIELD_K, FIELD_NU, FIELD_ALPHA) = range(12)
OPT_KS_UUROW, OPT_NUM_AGENTS, OPT_NUM_ACTIONS = 1, 2, 3
ABI_VERSION = 1


class MpdeConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("equation", C.c_int32), ("dtype", C.c_int32), ("device", C.c_int32),
        ("nenvs", C.c_int64),
        ("N", C.c_int32), ("M", C.c_int32), ("num_agents", C.c_int32), ("version", C.c_int32),
        ("stepper", C.c_int32), ("flags", C.c_int32), ("reward_mode", C.c_int32), ("team_lanes", C.c_int32),
        ("L", C.c_double), ("dt", C.c_double),
    ]


# name -> (restype, argtypes); the single source the symbol test checks against the header
_vp, _i32, _i64, _dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
SIGNATURES = {
    "mpde_create": (C.c_int, [C.POINTER(MpdeConfig), C.POINTER(_vp)]),
    "mpde_destroy": (C.c_int, [_vp]),
    "mpde_state_size": (_i64, [_vp]),
    "mpde_set_nu": (C.c_int, [_vp, _dp, _i64]),
    "mpde_set_basis": (C.c_int, [_vp, _i32, _dp]),
    "mpde_set_reward_mode": (C.c_int, [_vp, _i32]),
    "mpde_set_option": (C.c_int, [_vp, _i32, _i64]),
    "mpde_set_etdrk4": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "mpde_set_forcing": (C.c_int, [_vp, _dp, _i64]),
    "mpde_set_spectrum_ref": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "mpde_set_truth": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "mpde_set_history": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "mpde_reset_u": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mpde_reset_v": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mpde_step": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "mpde_reset_handoff": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mpde_reset_turbulence": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mpde_forcing_tables": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mpde_rng_last_error": (C.c_char_p, []),
    "mpde_fit_spline": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mpde_eval_spline_table": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _vp, _i32, _vp]),
    "mpde_step_host": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "mpde_step_host_packed": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "mpde_compute_sgs": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "mpde_get": (C.c_int, [_vp, _i32, _vp, _vp]),
    "mpde_set": (C.c_int, [_vp, _i32, _vp, _vp]),
    "mpde_launch_count": (_i64, [_vp]),
    "mpde_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "mpde_peer_free": (C.c_int, [_vp]),
    "mpde_peer_export": (C.c_int, [_vp, _vp]),
    "mpde_peer_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mpde_peer_close": (C.c_int, [_vp]),
    "mpde_peer_put": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp), C.c_size_t, C.POINTER(_vp), _i32, _i32, _i64, _vp, _vp]),
    "mpde_peer_wait": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp]),
    "mpde_peer_last_error": (C.c_char_p, []),
    "mpde_host_flag_alloc": (C.c_int, [_i32, C.POINTER(_vp)]),
    "mpde_host_flag_free": (C.c_int, [_vp]),
    "mpde_set_peer_sync": (C.c_int, [_vp, C.POINTER(_vp), _i32, _vp, _vp, _i32, _vp, _vp, _i64]),
    "mpde_step_fused": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "mpde_peer_join": (C.c_int, [_vp, _vp]),
    "mpde_set_peer_output": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(_vp), _i64, _vp, _vp]),
    "mpde_set_peer_local": (C.c_int, [_vp, _vp, _vp]),
    "mpde_set_peer_row_stores": (C.c_int, [_vp, C.c_int32]),
    "mpde_peer_signal_next": (C.c_int, [C.POINTER(_vp), _i32, _vp, _vp]),
    "mpde_peer_exchange_next": (C.c_int, [C.POINTER(_vp), _i32, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "mpde_peer_wait_next": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _vp]),
    "mpde_last_error": (C.c_char_p, []),
    "mpde_abi_version": (C.c_int, []),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: marlpde_b200 has no CPU/eager fallback. Build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C marlpde_b200/csrc).")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)           # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    if L.mpde_abi_version() != ABI_VERSION:
        raise LibraryMissing(f"{LIB_PATH} has ABI {L.mpde_abi_version()}, binding expects {ABI_VERSION}: rebuild")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("marlpde_b200: " + lib().mpde_last_error().decode())


def as_dp(a):
    """numpy float64 C-contiguous array -> double*"""
    return a.ctypes.data_as(_dp)
