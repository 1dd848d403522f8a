"""ctypes binding of include/sigma_b200.h -- the only way Python reaches the
CUDA path.  There is no fallback: if libsigma_b200.so is missing or no sm_100
device is usable, every compute call raises SigmaError."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIGB_LIB_VARIANT selects an experimental build (csrc/Makefile VARIANT=...); default: the product library
LIB_PATH = os.path.join(_HERE, "lib", "libsigma_b200%s.so" % os.environ.get("SIGB_LIB_VARIANT", ""))

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NONSQUARE, ERR_ISOLATED, ERR_COMM, ERR_UNSUPPORTED = range(8)
ROW, COL = 0, 1
FMT_CSR, FMT_CSC, FMT_ELLPACK = 1, 2, 3
UNIQUE_ID_BYTES = 128


class SigmaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"sigma_b200 status {status}: {message}")
        self.status = status
        self.message = message


_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_vp = C.c_void_p
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/sigma_b200.h one to one
PROTOTYPES = {
    "sigb_init": (C.c_int, [C.c_int]),
    "sigb_finalize": (C.c_int, []),
    "sigb_last_error": (C.c_char_p, []),
    "sigb_version": (C.c_char_p, []),
    "sigb_set_stream": (C.c_int, [_vp]),
    "sigb_synchronize": (C.c_int, []),
    "sigb_launch_count": (_i64, []),
    "sigb_dev_alloc": (C.c_int, [_i64, _pvp]),
    "sigb_dev_free": (C.c_int, [_vp]),
    "sigb_copy_h2d": (C.c_int, [_vp, _vp, _i64]),
    "sigb_copy_d2h": (C.c_int, [_vp, _vp, _i64]),
    "sigb_cs_graph_create": (C.c_int, [_i32, _i32, _vp, _vp, C.c_int, _pvp]),
    "sigb_ell_graph_create": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _pvp]),
    "sigb_cs_graph_build": (C.c_int, [_i32, _i32, _i64, _vp, _vp, C.c_int, C.c_int, _pvp]),
    "sigb_ell_graph_build": (C.c_int, [_i32, _i32, _i64, _vp, _vp, C.c_int, _pvp]),
    "sigb_graph_retain": (C.c_int, [_vp]),
    "sigb_graph_release": (C.c_int, [_vp]),
    "sigb_cs_graph_get_transpose": (C.c_int, [_vp, _vp, _vp]),
    "sigb_matrix_create": (C.c_int, [_vp, _pvp]),
    "sigb_matrix_set_values": (C.c_int, [_vp, _vp, _i64]),
    "sigb_matrix_destroy": (C.c_int, [_vp]),
    "sigb_matrix_get_dims": (C.c_int, [_vp, _pi32, _pi32, _pi64]),
    "sigb_matrix_get_transpose_values": (C.c_int, [_vp, _vp]),
    "sigb_matrix_copy": (C.c_int, [_vp, C.c_int, C.c_int, _pvp]),
    "sigb_matrix_get_format": (C.c_int, [_vp, C.POINTER(C.c_int), _pi32, _pi32, _pi64, _pi32]),
    "sigb_matrix_get_arrays": (C.c_int, [_vp, _vp, _vp, _vp]),
    "sigb_matrix_add_values": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "sigb_matvec": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "sigb_matvec_add": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "sigb_matvec_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "sigb_matvec_dot_dev": (C.c_int, [_vp, _vp, _vp, _pf64]),
    "sigb_operator_sum": (C.c_int, [_vp, _vp, _pvp]),
    "sigb_operator_product": (C.c_int, [_vp, _vp, _pvp]),
    "sigb_operator_adjoint": (C.c_int, [_vp, _pvp]),
    "sigb_composite_create": (C.c_int, [_i32, _i32, _vp, _vp, _pvp, _pvp]),
    "sigb_matrix_retain": (C.c_int, [_vp]),
    "sigb_cg_create": (C.c_int, [_f64, _pvp]),
    "sigb_bicgstab_create": (C.c_int, [_f64, _pvp]),
    "sigb_jacobi_create": (C.c_int, [_pvp]),
    "sigb_ldu_create": (C.c_int, [_pvp]),
    "sigb_ldu_get_sizes": (C.c_int, [_vp, _pi32, _pi64, _pi64, _pi32, _pi32]),
    "sigb_ldu_get_factors": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sigb_ldu_symbolic": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pi32, _vp, _vp, _pi32]),
    "sigb_solver_setup": (C.c_int, [_vp, _vp]),
    "sigb_solver_set_params": (C.c_int, [_vp, _f64]),
    "sigb_solver_set_max_iterations": (C.c_int, [_vp, _i64]),
    "sigb_solver_set_persistent": (C.c_int, [_vp, C.c_int]),
    "sigb_solver_set_strict_order": (C.c_int, [_vp, C.c_int]),
    "sigb_solver_solve": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sigb_solver_solve_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sigb_solver_get_info": (C.c_int, [_vp, _pi64, _pf64, C.POINTER(C.c_int)]),
    "sigb_solver_get_vector": (C.c_int, [_vp, C.c_char_p, _vp]),
    "sigb_solver_destroy": (C.c_int, [_vp]),
    "sigb_lanczos": (C.c_int, [_vp, _i32, _vp, C.c_uint64, _vp, _vp]),
    "sigb_lanczos_dev": (C.c_int, [_vp, _i32, _vp, C.c_uint64, _vp, _vp]),
    "sigb_eigensolve": (C.c_int, [_vp, _i32, _vp, C.c_uint64, _vp, _vp]),
    "sigb_generalized_lanczos": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, C.c_uint64, _vp, _vp]),
    "sigb_generalized_eigensolve": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, C.c_uint64, _vp, _vp]),
    "sigb_debug_row_tiles": (C.c_int, [_i32, _vp, _vp, _pi32]),
    "sigb_debug_row_tiles_dev": (C.c_int, [_i32, _vp, _vp, _pi32]),
    "sigb_debug_row_tiles_balanced": (C.c_int, [_i32, _vp, _i32, _vp, _pi32]),
    "sigb_debug_ldu_sweep_plan": (C.c_int, [_i32, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sigb_debug_cg_phase_cycles": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "sigb_debug_spmv_tile_cycles": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "sigb_partition_rows": (C.c_int, [_i32, _vp, _i32, _vp]),
    "sigb_halo_build": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _pi32, _vp]),
    "sigb_comm_unique_id": (C.c_int, [_vp]),
    "sigb_comm_create": (C.c_int, [_vp, C.c_int, C.c_int, _pvp]),
    "sigb_comm_destroy": (C.c_int, [_vp]),
    "sigb_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sigb_dist_csr_create": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _pvp]),
    "sigb_dist_get_halo": (C.c_int, [_vp, _pi32, _vp]),
    "sigb_mgpu_init": (C.c_int, [C.c_int]),
    "sigb_mgpu_finalize": (C.c_int, []),
    "sigb_mgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sigb_mgpu_csr_create": (C.c_int, [_i32, _vp, _vp, _pvp]),
}

_lib = None


def lib():
    """Load libsigma_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SigmaError(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                             " (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise SigmaError(status, lib().sigb_last_error().decode())


def ptr(a):
    """Raw address of a numpy array, torch tensor (host or device) or int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
