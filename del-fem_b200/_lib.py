"""ctypes loader of lib/libdelfem_b200.so (the C ABI declared in include/delfem_b200.h).

Fails loudly when the CUDA library has not been built: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdelfem_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "delfem_b200.h")


class DfbError(RuntimeError):
    """non-zero dfb_status (the reference would have panicked / returned Err)"""

    def __init__(self, status, message):
        super().__init__(f"dfb_status={status}: {message}")
        self.status = status


STATUS = {"OK": 0, "BAD_ARG": 1, "PATTERN_MISS": 2, "SINGULAR_DIAG": 3, "CUDA": 4, "NCCL": 5, "INDEX_OVERFLOW": 6,
          "NOT_POSITIVE": 7, "UNSUPPORTED": 8}


def build(force: bool = False) -> str:
    """compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)"""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
    subprocess.check_call(["make", "-C", _HERE, "-s", "-j8"])
    return LIB_PATH


def header_symbols():
    """every function name the public header declares"""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", txt)))


_lib = None
_HANDLE_GETTERS = ("dfb_bsr_pattern", "dfb_solver_sparse", "dfb_solver_ilu", "dfb_solver_r_vec", "dfb_solver_u_vec", "dfb_solver_ap_vec", "dfb_solver_p_vec")

_vp = C.c_void_p
_u64 = C.c_uint64
_i32 = C.c_int32
_f64 = C.c_double
_pp = C.POINTER(C.c_void_p)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C del-fem_b200` (or __graft_entry__.build()). "
            "del_fem_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.dfb_last_error_message.restype = C.c_char_p
    L.dfb_version.restype = C.c_char_p
    L.dfb_vec_device_ptr.restype = C.c_void_p
    # everything else returns dfb_status (int32)
    for name in header_symbols():
        fn = getattr(L, name)
        if name in ("dfb_last_error_message", "dfb_version", "dfb_vec_device_ptr"):
            continue
        if name in _HANDLE_GETTERS:
            fn.restype = C.c_void_p
            continue
        fn.restype = C.c_int32
    sig = {
        "dfb_ctx_create": [_i32, _pp],
        "dfb_ctx_destroy": [_vp],
        "dfb_ctx_sync": [_vp],
        "dfb_ctx_launch_count": [_vp, C.POINTER(_u64)],
        "dfb_ctx_set_option": [_vp, C.c_char_p, C.c_int64],
        "dfb_ctx_get_option": [_vp, C.c_char_p, C.POINTER(C.c_int64)],
        "dfb_timer_start": [_vp],
        "dfb_timer_stop": [_vp, C.POINTER(C.c_float)],
        "dfb_flush_l2": [_vp],
        "dfb_device_info": [_vp, C.POINTER(_i32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "dfb_host_register": [_vp, _vp, _u64],
        "dfb_host_unregister": [_vp, _vp],
        "dfb_profile_spmv": [_vp, C.POINTER(_f64), C.POINTER(_u64), _i32],
        "dfb_bcflag_create": [_vp, _vp, _u64, _i32, _pp],
        "dfb_bcflag_destroy": [_vp],
        "dfb_vec_set_fixed_dof_zero_dev": [_vp, _vp],
        "dfb_mesh_rotate_vertex_ids": [_vp, _u64],
        "dfb_bsr_set_fixed_dof_dev": [_vp, _f64, _vp],
        "dfb_vec_create": [_vp, _u64, _u64, _i32, _pp],
        "dfb_vec_destroy": [_vp],
        "dfb_vec_upload": [_vp, _vp],
        "dfb_vec_download": [_vp, _vp],
        "dfb_vec_set_zero": [_vp],
        "dfb_vec_copy": [_vp, _vp],
        "dfb_vec_dot": [_vp, _vp, C.POINTER(_f64)],
        "dfb_vec_add_scaled": [_vp, _f64, _vp],
        "dfb_vec_scale_and_add": [_vp, _f64, _vp],
        "dfb_vec_scale": [_vp, _f64],
        "dfb_vec_set_fixed_dof_zero": [_vp, _vp],
        "dfb_vec_fill_splitmix": [_vp, _u64, _u64, _f64, _f64],
        "dfb_vec_device_ptr": [_vp],
        "dfb_mesh_create": [_vp, _vp, _u64, _i32, _vp, _u64, _i32, _pp],
        "dfb_mesh_create_u32": [_vp, _vp, _u64, _i32, _vp, _u64, _i32, _pp],
        "dfb_mesh_create_kuhn": [_vp, _u64, _u64, _u64, _f64, _u64, _u64, _pp],
        "dfb_mesh_destroy": [_vp],
        "dfb_mesh_update_xyz": [_vp, _vp],
        "dfb_mesh_sizes": [_vp, C.POINTER(_u64), C.POINTER(_i32), C.POINTER(_u64), C.POINTER(_i32)],
        "dfb_mesh_download": [_vp, _vp, _vp],
        "dfb_pattern_create": [_vp, _vp, _vp, _u64, _i32, _pp],
        "dfb_pattern_create_u32": [_vp, _vp, _vp, _u64, _i32, _pp],
        "dfb_pattern_from_uniform_mesh": [_vp, _vp, _i32, _u64, _pp],
        "dfb_pattern_destroy": [_vp],
        "dfb_pattern_sizes": [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_i32)],
        "dfb_pattern_download": [_vp, _vp, _vp],
        "dfb_bsr_create": [_vp, _vp, _i32, _pp],
        "dfb_bsr_destroy": [_vp],
        "dfb_bsr_set_zero": [_vp],
        "dfb_bsr_upload": [_vp, _vp, _vp],
        "dfb_bsr_download": [_vp, _vp, _vp],
        "dfb_bsr_set_fixed_dof": [_vp, _f64, _vp],
        "dfb_bsr_mult_vec": [_vp, _vp, _f64, _f64, _vp],
        "dfb_bsr_mult_mat": [_vp, _vp, _f64, _f64, _vp],
        "dfb_bsr_mult_square_matrices": [_vp, _vp, _pp, _pp],
        "dfb_bsr_add_to_diagonal": [_vp, _f64, _vp],
        "dfb_bsr_add_scaled": [_vp, _f64, _vp],
        "dfb_bsr_merge_one": [_vp, _vp, _vp, _i32, _i32],
        "dfb_plan_create": [_vp, _vp, _vp, _i32, _pp],
        "dfb_plan_destroy": [_vp],
        "dfb_plan_download_elem2nz": [_vp, _vp],
        "dfb_plan_sizes": [_vp, C.POINTER(_u64), C.POINTER(_u64)],
        "dfb_assemble": [_vp, _i32, _vp, _f64, _f64, _vp],
        "dfb_emat_batch": [_vp, _i32, _vp, _f64, _f64, _vp, _vp, _vp, _vp],
        "dfb_assemble_from_emat": [_vp, _vp, _vp],
        "dfb_assemble_neohookean_tet": [_vp, _vp, _vp, _f64, _f64, _vp, _vp, C.POINTER(_f64)],
        "dfb_assemble_energy": [_vp, _i32, _vp, _vp, _f64, _f64, _vp, _vp, C.POINTER(_f64)],
        "dfb_lumped_mass": [_vp, _vp, _f64, _vp],
        "dfb_emat_one": [_vp, _i32, _vp, _f64, _f64, _vp],
        "dfb_malloc": [_vp, _u64, _pp],
        "dfb_free": [_vp, _vp],
        "dfb_memcpy_d2h": [_vp, _vp, _vp, _u64],
        "dfb_memcpy_h2d": [_vp, _vp, _vp, _u64],
        "dfb_precond_create": [_vp, _i32, _vp, _pp],
        "dfb_precond_create_iluk": [_vp, _vp, _u64, _pp],
        "dfb_precond_update": [_vp, _vp],
        "dfb_precond_apply": [_vp, _vp],
        "dfb_precond_destroy": [_vp],
        "dfb_cg": [_vp, _vp, _vp, _vp, _vp, _f64, _u64, _f64, _i32, _vp, _u64, C.POINTER(_u64)],
        "dfb_pcg": [_vp, _vp, _vp, _vp, _vp, _vp, _f64, _u64, _f64, _vp, _u64, C.POINTER(_u64)],
        "dfb_comm_unique_id": [_vp],
        "dfb_ctx_comm_init": [_vp, _vp, _i32, _i32],
        "dfb_ctx_peer_export": [_vp, _vp],
        "dfb_ctx_peer_import": [_vp, _vp, _i32],
        "dfb_halo_create": [_vp, _i32, _vp, _vp, _vp, _vp, _pp],
        "dfb_halo_destroy": [_vp],
        "dfb_halo_exchange": [_vp, _vp],
        "dfb_halo_set_remote": [_vp, _vp],
        "dfb_bsr_set_halo": [_vp, _vp, _u64, _u64],
        "dfb_bsr_pattern": [_vp],
        "dfb_bsr_pack": [_vp],
        "dfb_ctx_phase_times": [_vp, _vp, _vp, _i32],
        "dfb_ctx_last_history_sq": [_vp, _vp, _u64, C.POINTER(_u64)],
        "dfb_solver_new": [_vp, _i32, _pp],
        "dfb_solver_initialize": [_vp, _vp, _u64, _vp, _u64],
        "dfb_solver_set_preconditioner": [_vp, _i32, _u64],
        "dfb_solver_begin_merge": [_vp],
        "dfb_solver_end_merge": [_vp],
        "dfb_solver_solve_cg": [_vp],
        "dfb_solver_solve_pcg": [_vp],
        "dfb_solver_clone": [_vp, _pp],
        "dfb_solver_destroy": [_vp],
        "dfb_solver_sparse": [_vp],
        "dfb_solver_ilu": [_vp],
        "dfb_solver_r_vec": [_vp],
        "dfb_solver_u_vec": [_vp],
        "dfb_solver_ap_vec": [_vp],
        "dfb_solver_p_vec": [_vp],
        "dfb_solver_set_params": [_vp, _f64, _u64],
        "dfb_solver_get_params": [_vp, C.POINTER(_f64), C.POINTER(_u64)],
        "dfb_solver_conv": [_vp, _vp, _u64, C.POINTER(_u64)],
    }
    for name, argtypes in sig.items():
        getattr(L, name).argtypes = argtypes
    _lib = L
    return L


def check(status: int):
    if status != 0:
        raise DfbError(status, lib().dfb_last_error_message().decode())
