"""ctypes binding of libtnl_b200.so (include/tnl_b200.h).  There is no CPU fallback: if the library is
missing it is built with nvcc; if no GPU is visible `Context()` raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtnl_b200.so")


class TnlError(RuntimeError):
    pass


class tnl_index_t(C.Structure):
    _fields_ = [("nsect", C.c_int32), ("dir", C.c_int32), ("dims", C.POINTER(C.c_int32)), ("qns", C.POINTER(C.c_int32))]


_lib = None

# name -> argtypes (restype is always int except tnl_last_error)
_P = C.c_void_p
_SIGS = {
    "tnl_ctx_create": [C.c_int, C.POINTER(_P)],
    "tnl_ctx_destroy": [_P],
    "tnl_get_counters": [_P, C.POINTER(C.c_double)],
    "tnl_reset_counters": [_P],
    "tnl_ctx_sync": [_P],
    "tnl_ctx_reserve": [_P, C.c_int64],
    "tnl_timer_start": [_P, C.c_int32],
    "tnl_timer_stop": [_P, C.c_int32, C.POINTER(C.c_double)],
    "tnl_comm_unique_id": [C.c_char_p],
    "tnl_comm_init": [_P, C.c_char_p, C.c_int32, C.c_int32],
    "tnl_comm_destroy": [_P],
    "tnl_comm_set_sharding": [_P, C.c_int32],
    "tnl_comm_bench": [_P, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_double)],
    "tnl_shard_range": [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "tnl_gemm_selftest": [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                          C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "tnl_profile_gemm": [_P, C.c_int32],
    "tnl_profile_read": [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "tnl_profile_categories": [_P, C.POINTER(C.c_double)],
    "tnl_profile_collectives": [_P, C.POINTER(C.c_double)],
    "tnl_tensor_import": [_P, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64, _P, _P, _P, C.c_int32, C.POINTER(_P)],
    "tnl_tensor_import_c128": [_P, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64, _P, _P, _P, C.c_int32,
                               C.POINTER(_P)],
    "tnl_tensor_is_complex": [_P, C.POINTER(C.c_int32)],
    "tnl_tensor_promote": [_P],
    "tnl_vec_dot_c": [_P, _P, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "tnl_tensor_create": [_P, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int32, C.POINTER(_P)],
    "tnl_tensor_free": [_P],
    "tnl_tensor_copy": [_P, C.POINTER(_P)],
    "tnl_tensor_rank": [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "tnl_tensor_index": [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P, C.c_int32],
    "tnl_tensor_export_size": [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "tnl_tensor_export": [_P, _P, _P, _P],
    "tnl_tensor_fill_random": [_P, C.c_uint64],
    "tnl_tensor_scale_index": [_P, C.c_int32, _P],
    "tnl_vec_dot": [_P, _P, C.POINTER(C.c_double)],
    "tnl_vec_norm": [_P, C.POINTER(C.c_double)],
    "tnl_vec_scale": [_P, C.c_double],
    "tnl_vec_axpy": [_P, _P, C.c_double],
    "tnl_env_create": [_P, C.c_int32, C.POINTER(_P)],
    "tnl_env_destroy": [_P],
    "tnl_env_set_site_op": [_P, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64, _P, _P, _P],
    "tnl_env_update_site_op": [_P, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64, _P, _P, _P],
    "tnl_env_set_site_op_term": [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64, _P, _P, _P],
    "tnl_env_cm_set_term": [_P, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(tnl_index_t), C.c_int64,
                            _P, _P, _P],
    "tnl_env_add_penalty": [_P, C.c_double, C.c_int32, C.POINTER(_P)],
    "tnl_env_set_state": [_P, C.c_int32, _P],
    "tnl_env_get_state": [_P, C.c_int32, C.POINTER(_P)],
    "tnl_env_set_nsite": [_P, C.c_int32],
    "tnl_env_position": [_P, C.c_int32],
    "tnl_env_move_center": [_P, C.c_int32, C.c_int32],
    "tnl_env_make_phi": [_P, C.c_int32, C.POINTER(_P)],
    "tnl_env_apply_flops": [_P, C.POINTER(C.c_double)],
    "tnl_heff_apply": [_P, _P, C.POINTER(_P)],
    "tnl_eigsolve_lanczos": [_P, _P, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)],
    "tnl_expectation": [_P, _P, C.POINTER(C.c_double)],
    "tnl_svd_split": [_P, C.c_int32, _P, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_int32, C.c_int32,
                      C.POINTER(C.c_double), _P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(_P)],
    "tnl_env_absorb_bond": [_P, C.c_int32, C.c_int32, _P],
    "tnl_exponentiate": [_P, _P, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)],
    "tnl_tensor_permute": [_P, _P, C.c_int32, C.POINTER(_P)],
    "tnl_tensor_dag": [_P, C.POINTER(_P)],
    "tnl_tensor_nrow": [_P, C.POINTER(C.c_int32)],
    "tnl_tensor_contract": [_P, _P, C.c_int32, _P, _P, C.c_int32, C.POINTER(_P), _P, C.POINTER(C.c_int32)],
    "tnl_tensor_directsum": [_P, C.c_int32, _P, C.c_int32, C.POINTER(_P)],
    "tnl_tensor_factorize": [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_int32, C.POINTER(_P),
                             C.POINTER(_P), C.POINTER(C.c_double), _P, C.c_int64, C.POINTER(C.c_int64)],
    "tnl_sumop_create": [_P, C.c_int32, _P, C.POINTER(_P)],
    "tnl_sumop_destroy": [_P],
    "tnl_sumop_add_term": [_P, C.c_int32, C.POINTER(_P), _P],
    "tnl_sumop_set_relabel": [_P, C.c_int32, _P, _P],
    "tnl_sumop_add_projector": [_P, _P, C.c_double],
    "tnl_sumop_apply": [_P, _P, C.POINTER(_P)],
    "tnl_sumop_apply_flops": [_P, C.POINTER(C.c_double)],
    "tnl_sumop_eigsolve": [_P, _P, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)],
    "tnl_sumop_exponentiate": [_P, _P, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                               C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)],
    "tnl_replacebond": [_P, C.c_int32, _P, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int32,
                        C.c_int32, C.POINTER(C.c_double), _P, C.c_int64, C.POINTER(C.c_int64)],
}
EXPORTED = sorted(list(_SIGS) + ["tnl_last_error"])


def so_path() -> str:
    return _SO


def load(build_if_missing: bool = True):
    """Load the CUDA library.  Raises if it cannot be built/loaded -- never falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_tnl_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    if not os.path.exists(_SO):
        raise TnlError(f"{_SO} is missing: the CUDA extension is required (no CPU fallback)")
    lib = C.CDLL(_SO, mode=C.RTLD_LOCAL)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.tnl_last_error.argtypes = [_P]
    lib.tnl_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, ctx=None):
    if rc != 0:
        msg = load().tnl_last_error(ctx)
        raise TnlError(f"tnl_b200 error {rc}: {msg.decode() if msg else '?'}")
