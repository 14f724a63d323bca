"""ctypes binding of libabr.so (include/abr.h).  There is no fallback: if the
CUDA library is missing or no device is present, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ABR_LIB_PATH: another build of the same library (tuning experiments); still no fallback
LIB_PATH = os.environ.get("ABR_LIB_PATH") or os.path.join(_HERE, "lib", "libabr.so")
MAX_VARS, MAX_PARAMS, MAX_D = 4, 8, 3

_lib = None


class AbrError(RuntimeError):
    pass


class KernelDesc(C.Structure):
    _fields_ = [
        ("kernel_id", C.c_int32),
        ("block_rows", C.c_int32),
        ("block_cols", C.c_int32),
        ("reserved", C.c_int32),
        ("params", C.c_double * MAX_PARAMS),
        ("row_vars", C.c_void_p * MAX_VARS),
        ("col_vars", C.c_void_p * MAX_VARS),
    ]


# name -> (restype, argtypes); also used by the symbol-export test
SIGNATURES = {
    "abr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "abr_destroy": (C.c_int, [C.c_void_p]),
    "abr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "abr_synchronize": (C.c_int, [C.c_void_p]),
    "abr_check_async": (C.c_int, [C.c_void_p]),
    "abr_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "abr_last_error_string": (C.c_char_p, [C.c_void_p]),
    "abr_version": (C.c_char_p, []),
    "abr_domain_set": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "abr_domain_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "abr_domain_force_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abr_grid_for": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, C.c_void_p]),
    "abr_domain_set_window": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "abr_celllist_adopt_sorted": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_celllist_patch_ghosts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abr_slab_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "abr_slab_layers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "abr_celllist_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "abr_celllist_get": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "abr_gather_columns": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_update_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "abr_query_set_particles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_sparse_matvec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(KernelDesc), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "abr_sparse_assemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(KernelDesc), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "abr_sparse_coeff": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(KernelDesc), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "abr_bucket_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "abr_fast_bucket_search_counts": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "abr_id_map_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_id_map_get": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "abr_id_find": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "abr_pair_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "abr_distance_search_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "abr_distance_search_stats_scaled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abr_distance_search_stats_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "abr_last_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64 * 4)]),
    "abr_sparse_matvec_custom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "abr_probe_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "abr_malloc": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]),
    "abr_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "abr_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "abr_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "abr_host_alloc_pinned": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "abr_host_free_pinned": (C.c_int, [C.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AbrError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C aboria_b200/csrc).  aboria_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("ABR_LIB_PATH") and not hasattr(L, name):
                continue  # an older build of the library in a tuning experiment: newer entry points are simply absent
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(h, rc):
    if rc != 0:
        msg = lib().abr_last_error_string(h)
        raise AbrError(f"libabr error {rc}: {msg.decode() if msg else ''}")
