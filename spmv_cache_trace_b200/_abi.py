"""ctypes declarations for libspmvb200.so -- one entry per function in include/spmv_b200.h.

The library is loaded on first use.  There is no fallback of any kind: if the
shared object is missing (not built) the import of the symbol table raises, and if
no CUDA device is usable every compute entry point returns SPMVB200_ERR_CUDA, which
the Python layer turns into `matrix_error`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPMVB200_LIB") or os.path.join(HERE, "lib", "libspmvb200.so")  # SPMVB200_LIB: an experiment build

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)


class Info(C.Structure):
    """spmvb200_info"""
    _fields_ = [
        ("format", C.c_int32), ("coo_mode", C.c_int32),
        ("rows", C.c_int64), ("columns", C.c_int64),
        ("num_entries", C.c_int64), ("stored_entries", C.c_int64),
        ("row_alignment", C.c_int64), ("ell_row_length", C.c_int64),
        ("num_ell_entries", C.c_int64), ("num_coo_entries", C.c_int64),
        ("skip_padding", C.c_int32), ("offsets_64bit", C.c_int32),
        ("matrix_size", C.c_int64), ("x_size", C.c_int64), ("y_size", C.c_int64),
        ("device_bytes", C.c_int64), ("row_offset", C.c_int64),
    ]


class CacheConfig(C.Structure):
    """spmvb200_cache_config"""
    _fields_ = [
        ("cache_bytes", C.c_int64), ("line_bytes", C.c_int32), ("parts", C.c_int32),
        ("starts", i64p), ("shared", C.c_int32), ("warmup", C.c_int32),
        ("page_bytes", C.c_int32), ("stream_bypass", C.c_int32),
    ]


class CacheMisses(C.Structure):
    """spmvb200_cache_misses"""
    _fields_ = [(n, C.c_int64) for n in (
        "references", "misses_index", "misses_column_index", "misses_value", "misses_x_local",
        "misses_x_remote", "misses_y_local", "misses_y_remote", "x_references", "x_remote_references")]


class DistInfo(C.Structure):
    """spmvb200_dist_info_t"""
    _fields_ = [
        ("rank", C.c_int32), ("nranks", C.c_int32), ("exchange", C.c_int32), ("n_blocks", C.c_int32),
        ("n_sends", C.c_int32), ("n_recvs", C.c_int32),
        ("recv_bytes_per_step", C.c_int64), ("send_bytes_per_step", C.c_int64),
        ("rows", C.c_int64), ("row_begin", C.c_int64), ("num_entries", C.c_int64), ("interior_rows", C.c_int64),
        ("device_bytes", C.c_int64), ("launches_per_step", C.c_int64), ("steps_done", C.c_int64), ("halo_push", C.c_int64),
    ]


_ccp = C.POINTER(CacheConfig)
_cmp = C.POINTER(CacheMisses)

# name -> (restype, argtypes); the exported symbol set the tests check against the header
SIGNATURES = {
    "spmvb200_last_error": (C.c_char_p, []),
    "spmvb200_version": (C.c_int, []),
    "spmvb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "spmvb200_set_device": (C.c_int, [C.c_int]),
    "spmvb200_device_props": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), i64p, i64p,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spmvb200_launch_count": (C.c_int64, []),
    "spmvb200_set_global_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "spmvb200_mm_parse": (C.c_int, [C.c_char_p, C.c_size_t, vpp]),
    "spmvb200_mm_load": (C.c_int, [C.c_char_p, vpp]),
    "spmvb200_mm_from_entries": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, i32p, i32p, f64p, vpp]),
    "spmvb200_mm_info": (C.c_int, [vp, i32p, i32p, i32p, i32p, i32p, i32p]),
    "spmvb200_mm_entries": (C.c_int, [vp, C.POINTER(i32p), C.POINTER(i32p), C.POINTER(f64p)]),
    "spmvb200_mm_max_row_length": (C.c_int, [vp, i32p]),
    "spmvb200_mm_row_lengths": (C.c_int, [vp, i32p]),
    "spmvb200_mm_sort_row_major": (C.c_int, [vp]),
    "spmvb200_mm_sort_column_major": (C.c_int, [vp]),
    "spmvb200_mm_free": (None, [vp]),
    "spmvb200_csr_from_mm": (C.c_int, [vp, C.c_int32, vpp]),
    "spmvb200_coo_from_mm": (C.c_int, [vp, C.c_int32, vpp]),
    "spmvb200_ell_from_mm": (C.c_int, [vp, C.c_int32, vpp]),
    "spmvb200_hyb_from_mm": (C.c_int, [vp, C.c_int32, vpp]),
    "spmvb200_csr_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, i32p, i32p, f64p, vpp]),
    "spmvb200_csr_create64": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, i64p, i32p, f64p, vpp]),
    "spmvb200_coo_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, i32p, i32p, f64p, C.c_int32, vpp]),
    "spmvb200_ell_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, f64p, C.c_int32, vpp]),
    "spmvb200_hyb_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, f64p, C.c_int32,
                                      C.c_int32, i32p, i32p, f64p, vpp]),
    "spmvb200_gen_stencil": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_int32, vpp]),
    "spmvb200_gen_rmat": (C.c_int, [C.c_int32, C.c_int32, C.c_uint64, C.c_double, C.c_double, C.c_double,
                                    C.c_int64, C.c_int64, C.c_int32, C.c_int32, vpp]),
    "spmvb200_convert": (C.c_int, [vp, C.c_int32, C.c_int32, vpp]),
    "spmvb200_matrix_info": (C.c_int, [vp, C.POINTER(Info)]),
    "spmvb200_csr_export": (C.c_int, [vp, i64p, i32p, f64p]),
    "spmvb200_coo_export": (C.c_int, [vp, i32p, i32p, f64p]),
    "spmvb200_ell_export": (C.c_int, [vp, i32p, f64p]),
    "spmvb200_hyb_export": (C.c_int, [vp, i32p, f64p, i32p, i32p, f64p]),
    "spmvb200_set_x": (C.c_int, [vp, f64p]),
    "spmvb200_set_y": (C.c_int, [vp, f64p]),
    "spmvb200_get_x": (C.c_int, [vp, f64p]),
    "spmvb200_get_y": (C.c_int, [vp, f64p]),
    "spmvb200_fill_x": (C.c_int, [vp, C.c_double]),
    "spmvb200_fill_y": (C.c_int, [vp, C.c_double]),
    "spmvb200_x_device": (C.c_int, [vp, vpp]),
    "spmvb200_y_device": (C.c_int, [vp, vpp]),
    "spmvb200_bind_x": (C.c_int, [vp, vp]),
    "spmvb200_bind_y": (C.c_int, [vp, vp]),
    "spmvb200_set_stream": (C.c_int, [vp, vp]),
    "spmvb200_host_alloc": (C.c_int, [C.c_size_t, vpp]),
    "spmvb200_host_free": (C.c_int, [vp]),
    "spmvb200_prepare": (C.c_int, [vp]),
    "spmvb200_set_alpha": (C.c_int, [vp, C.c_double]),
    "spmvb200_spmv": (C.c_int, [vp]),
    "spmvb200_sync": (C.c_int, [vp]),
    "spmvb200_spmv_host": (C.c_int, [vp, f64p, f64p]),
    "spmvb200_time": (C.c_int, [vp, C.c_int, C.c_int, f32p]),
    "spmvb200_time_copy": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, f32p]),
    "spmvb200_time_rotating": (C.c_int, [vpp, C.c_int, C.c_int, C.c_int, f32p, f32p]),
    "spmvb200_time_host_rotating": (C.c_int, [vpp, C.c_int, C.POINTER(f64p), C.POINTER(f64p), C.c_int, C.c_int,
                                              f32p]),
    "spmvb200_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "spmvb200_get_option": (C.c_int, [vp, C.c_char_p, i64p]),
    "spmvb200_kernel_name": (C.c_char_p, [vp]),
    "spmvb200_destroy": (C.c_int, [vp]),
    "spmvb200_partition_rows_ref": (C.c_int, [C.c_int64, C.c_int32, i64p]),
    "spmvb200_partition_rows_nnz": (C.c_int, [vp, C.c_int32, i64p]),
    "spmvb200_partition_rows_weighted": (C.c_int, [vp, C.c_int32, C.c_int64, i64p]),
    "spmvb200_csr_row_block": (C.c_int, [vp, C.c_int64, C.c_int64, vpp]),
    "spmvb200_csr_column_split": (C.c_int, [vp, C.c_int64, C.c_int64, vpp, vpp]),
    "spmvb200_csr_column_span": (C.c_int, [vp, C.c_int64, C.c_int64, i64p, i64p, i64p, i64p]),
    "spmvb200_mm_order_rcm": (C.c_int, [vp, i32p]),
    "spmvb200_mm_order_gp": (C.c_int, [vp, C.c_int32, i32p]),
    "spmvb200_mm_order_gp_kway": (C.c_int, [vp, C.c_int32, i32p]),
    "spmvb200_mm_partition_kway": (C.c_int, [vp, C.c_int32, C.c_int32, i32p, C.POINTER(C.c_int64)]),
    "spmvb200_order_from_parts": (C.c_int, [C.c_int32, C.c_int32, i32p, i32p]),
    "spmvb200_mm_permute": (C.c_int, [vp, i32p]),
    "spmvb200_cache_trace_csr": (C.c_int, [C.c_int64, C.c_int64, i64p, i32p, _ccp, _cmp]),
    "spmvb200_cache_trace_ell": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, i32p, _ccp, _cmp]),
    "spmvb200_cache_trace_coo": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, i32p, i32p, _ccp, _cmp]),
    "spmvb200_cache_trace": (C.c_int, [vp, _ccp, _cmp]),
    "spmvb200_comm_create_local": (C.c_int, [C.c_int, C.POINTER(C.c_int), vpp]),
    "spmvb200_comm_unique_id": (C.c_int, [vp]),
    "spmvb200_comm_create_nccl": (C.c_int, [vp, C.c_int, C.c_int, vpp]),
    "spmvb200_comm_rank": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spmvb200_comm_barrier": (C.c_int, [vp]),
    "spmvb200_comm_allreduce": (C.c_int, [vp, f64p, C.c_int]),
    "spmvb200_comm_destroy": (C.c_int, [vp]),
    "spmvb200_exchange_plan": (C.c_int, [C.c_int32, i64p, i64p, i64p, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i64p,
                                         i32p, i64p, i64p]),
    "spmvb200_dist_create": (C.c_int, [vp, vp, i64p, C.c_int32, C.c_int32, C.c_int32, vpp]),
    "spmvb200_dist_set_x": (C.c_int, [vp, f64p]),
    "spmvb200_dist_get_x": (C.c_int, [vp, f64p]),
    "spmvb200_dist_x_device": (C.c_int, [vp, vpp]),
    "spmvb200_dist_step": (C.c_int, [vp, C.c_double]),
    "spmvb200_dist_sync": (C.c_int, [vp]),
    "spmvb200_dist_time": (C.c_int, [vpp, C.c_int, C.c_int, C.c_int, C.c_double, f32p]),
    "spmvb200_dist_run_host": (C.c_int, [vp, C.c_int, C.POINTER(f64p), C.POINTER(f64p), C.c_double, f32p]),
    "spmvb200_dist_info": (C.c_int, [vp, C.POINTER(DistInfo)]),
    "spmvb200_dist_block": (C.c_int, [vp, C.c_int32, i64p, i64p, i32p, vpp]),
    "spmvb200_dist_destroy": (C.c_int, [vp]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m spmv_cache_trace_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
