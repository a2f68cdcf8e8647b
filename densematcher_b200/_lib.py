"""ctypes binding of libdm_b200.so (include/dm_b200.h).

This is the whole "FFI": the library is a plain C-ABI object taking device pointers, and this
module is the stub a maintainer of the reference would add next to
``densematcher/pyFM/spectral/nn_utils.py``.  There is NO fallback: if the
shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdm_b200.so")

# flags (keep in sync with include/dm_b200.h)
DM_I64_OUT = 1 << 0
DM_NO_RECHECK = 1 << 1
DM_ENGINE_FFMA = 1 << 2
DM_ENGINE_TC = 1 << 3
DM_RECHECK_ALL = 1 << 4
DM_SKIP_PREP = 1 << 5
DM_SKIP_FINISH = 1 << 6
DM_F64_GEMM = 1 << 7
DM_FAST_FM = 1 << 8
DM_FAST_LOSS = 1 << 9
DM_POLAR_JACOBI = 1 << 9
SCALE_NONE, SCALE_ARRAY, SCALE_INVNORM = 0, 1, 2
BIAS_NONE, BIAS_ARRAY, BIAS_NEG_HALF_SQNORM = 0, 1, 2

c_i64 = C.c_int64
c_int = C.c_int
c_vp = C.c_void_p
c_sz = C.c_size_t
c_dbl = C.c_double


class NNEpi(C.Structure):
    """struct dm_nn_epi"""
    _fields_ = [("scale_mode", C.c_int32), ("bias_mode", C.c_int32), ("scale", c_vp), ("bias", c_vp),
                ("out", c_vp)]


# name -> (restype, argtypes); every symbol include/dm_b200.h declares
SIGNATURES = {
    "dm_last_error": (C.c_char_p, []),
    "dm_version": (c_int, []),
    "dm_build_info": (C.c_char_p, []),
    "dm_nn_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dm_nn_argmax_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int,
                                 C.POINTER(NNEpi), c_int, C.POINTER(NNEpi), c_int, c_int, c_vp, c_sz, c_vp]),
    "dm_nn_f64_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dm_nn_argmax_f64": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int,
                                 C.POINTER(NNEpi), c_int, C.POINTER(NNEpi), c_int, c_int, c_vp, c_sz, c_vp]),
    "dm_nn_read_stats": (c_int, [c_vp, C.POINTER(c_i64), c_vp]),
    "dm_nn_debug_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "dm_nn_debug_scores_f32": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_int, c_vp,
                                       c_sz, c_vp]),
    "dm_knn_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "dm_knn_f64": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dm_match_dist_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_int, c_vp]),
    "dm_project_workspace_bytes": (c_sz, [c_int, c_i64, c_int, c_int, c_int]),
    "dm_project": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp,
                           c_sz, c_vp]),
    "dm_project_ex": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_int,
                              c_vp, c_sz, c_vp]),
    "dm_fmap_solve_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "dm_fmap_solve": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_vp, c_vp,
                              c_sz, c_vp]),
    "dm_fmap_c00": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "dm_fmap_solve_read_status": (c_int, [c_vp, C.POINTER(c_int), c_vp]),
    "dm_match_pairs_read_status": (c_int, [c_vp, C.POINTER(c_int), c_vp]),
    "dm_icp_read_status": (c_int, [c_vp, C.POINTER(c_int), c_vp]),
    "dm_fm_to_p2p_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int]),
    "dm_fm_to_p2p": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int,
                             c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_mapped_indicator_workspace_bytes": (c_sz, [c_int, c_int]),
    "dm_mapped_indicator": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_i64,
                                    c_vp, c_sz, c_vp]),
    "dm_mapped_indicators_workspace_bytes": (c_sz, [c_i64, c_int]),
    "dm_mapped_indicators": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_int, c_vp,
                                     c_int, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "dm_p2p_to_fm_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "dm_p2p_to_fm": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_int, c_int, c_int, c_vp,
                             c_int, c_vp, c_sz, c_vp]),
    "dm_spd_solve_workspace_bytes": (c_sz, [c_int, c_int]),
    "dm_spd_solve": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "dm_zoomout_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_int]),
    "dm_zoomout": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64,
                           c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_dense_energy_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int]),
    "dm_dense_energy": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int,
                                c_vp, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dm_dense_energy_ex": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int,
                                   c_vp, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_bmm_nt_f64": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "dm_match_pairs_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int]),
    "dm_match_pairs": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                               c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_int, c_int, c_dbl, c_dbl,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_bank_state_bytes": (c_sz, [c_int, c_i64, c_int, c_int]),
    "dm_bank_prepare_workspace_bytes": (c_sz, [c_int, c_i64, c_int, c_int, c_int]),
    "dm_bank_prepare": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_i64,
                                c_i64, c_vp, c_sz, c_vp, c_sz, c_vp]),
    "dm_match_bank_pairs_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int]),
    "dm_match_bank_pairs": (c_int, [c_vp, c_sz, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int,
                                    c_vp, c_vp,
                                    c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_int, c_int, c_dbl, c_dbl,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_match_bank_pairs_read_status": (c_int, [c_vp, c_int, c_int, c_int, C.POINTER(c_int), c_vp]),
    "dm_polar_factor_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "dm_polar_factor": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_lap_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_i64]),
    "dm_lap_solve": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_precise_map_workspace_bytes": (c_sz, [c_int, c_i64, c_int, c_i64, c_i64]),
    "dm_precise_map": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_int,
                               c_int, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    "dm_lbo_eigs_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "dm_lbo_eigs": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_dbl, c_int, c_int, c_vp, c_vp, c_i64, c_vp,
                            c_vp, c_vp, c_sz, c_vp]),
    "dm_sym_eig_workspace_bytes": (c_sz, [c_int, c_int]),
    "dm_sym_eig": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dm_from_basis": (c_int, [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp]),
    "dm_spectral_diffusion_workspace_bytes": (c_sz, [c_int, c_i64, c_int, c_int, c_int]),
    "dm_spectral_diffusion": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int,
                                      c_int, c_vp, c_i64, c_int, c_vp, c_sz, c_vp]),
    "dm_fps_workspace_bytes": (c_sz, [c_i64]),
    "dm_fps": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_sz, c_vp]),
    "dm_icp_workspace_bytes": (c_sz, [c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int]),
    "dm_icp": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_int,
                       c_int, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
}

_lib = None


class DMError(RuntimeError):
    pass


def load():
    """Loads libdm_b200.so (once).  Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -m densematcher_b200.build` (needs nvcc, no GPU required). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library is stale: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dm_last_error()
        raise DMError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) pointer of a torch tensor, or None."""
    return None if t is None else C.c_void_p(t.data_ptr())
