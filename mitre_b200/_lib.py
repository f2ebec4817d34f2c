"""ctypes binding of libmitre_b200.so (the C ABI in include/mitre_b200.h).

There is no fallback: if the library is missing or fails to load, importing this module's
`lib()` raises.  Build it with `python -m mitre_b200.build` (nvcc, sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmitre_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_double = ctypes.c_double
c_size_t = ctypes.c_size_t

MITRE_OK = 0
MITRE_ERR_NOT_PD = -4
STATUS_BASE_NOT_PD = 1
STATUS_BAD_SCHUR = 2
STATUS_COV_NOT_PD = 4
STATUS_COMM_TIMEOUT = 8
MAX_P = 24

# name -> (restype, argtypes); every symbol include/mitre_b200.h declares
SIGNATURES = {
    "mitre_version": (ctypes.c_char_p, []),
    "mitre_error_string": (ctypes.c_char_p, [c_int]),
    "mitre_device_info": (c_int, [ctypes.POINTER(c_int)] * 3 + [ctypes.POINTER(c_size_t)]),
    "mitre_probe_fp64": (c_int, [c_void_p, c_int, c_void_p, ctypes.POINTER(c_double)]),
    "mitre_pack_rows_f64": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "mitre_pack_rows_u8": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "mitre_unpack_rows_u8": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "mitre_score_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mitre_set_large_table_min_rows": (c_int, [c_i64]),
    "mitre_set_split_tables": (c_int, [c_int]),
    "mitre_set_groups_warp": (c_int, [c_int]),
    "mitre_set_wide_tables": (c_int, [c_int]),
    "mitre_set_dedupe": (c_int, [c_int]),
    "mitre_set_dedupe_table_bits": (c_int, [c_int]),
    "mitre_score_all": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p,
                                c_void_p]),
    "mitre_score_groups": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p,
                                   c_void_p, c_void_p]),
    "mitre_score_all_stats": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "mitre_score_all_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int,
                                     c_int, c_i64, c_void_p]),
    "mitre_likelihood_z_given_X_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                                 c_void_p, c_int, c_int, c_double, c_void_p,
                                                 c_void_p]),
    "mitre_mc_logpdf": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p]),
    "mitre_window_ranges": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p]),
    "mitre_detector_values": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_void_p]),
    "mitre_detector_thresholds": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "mitre_scan_workspace_bytes": (c_size_t, [c_i64]),
    "mitre_detector_row_offsets": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p,
                                           c_size_t]),
    "mitre_detector_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                    c_int, c_void_p]),
    "mitre_lexsort_workspace_bytes": (c_size_t, [c_i64, c_int]),
    "mitre_lexsort_rows": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_size_t]),
    "mitre_lexsort_padded": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_i64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t]),
    "mitre_lexsort_check": (c_int, [c_void_p, c_void_p, c_i64, c_int]),
    "mitre_set_packed_sort": (c_int, [c_int]),
    "mitre_window_lists": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_i64, c_i64] +
                           [c_void_p] * 7),
    "mitre_detectors_fused": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int] + [c_void_p] * 7 +
                              [c_i64] + [c_void_p] * 4 + [c_int, c_int, c_i64] + [c_void_p] * 7),
    "mitre_window_tbar": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mitre_values_walk_smem_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mitre_detector_values_walk": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_void_p, c_int] + [c_void_p] * 7 + [c_int, c_int, c_int, c_void_p,
                                                                                 c_void_p]),
    "mitre_detectors_walk": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_int] + [c_void_p] * 7 + [c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                                                 c_i64] + [c_void_p] * 7),
    "mitre_detector_sort_emit": (c_int, [c_void_p, c_void_p, c_i64, c_int] + [c_void_p] * 6),
    "mitre_slope_guard": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "mitre_set_detector_split": (c_int, [c_int]),
    "mitre_set_walk_vpt": (c_int, [c_int]),
    "mitre_set_host_chunk_bytes": (c_int, [c_i64]),
    "mitre_set_walk_tuning": (c_int, [c_int, c_int, c_int]),
    "mitre_walk_vars_per_block": (c_int, [c_int]),
    "mitre_set_sort_emit_rolled": (c_int, [c_int]),
    "mitre_selftest_divide": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "mitre_densify_perm": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "mitre_row_groups": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "mitre_row_prior": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_double, c_void_p]),
    "mitre_gibbs_omega_beta": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_int,
                                       ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mitre_pg_draw": (c_int, [c_void_p, c_void_p, c_i64, ctypes.c_uint64, c_void_p]),
    "mitre_draw_workspace_bytes": (c_size_t, [c_i64]),
    "mitre_weight_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p,
                                   c_size_t]),
    "mitre_draw_index": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_double, c_int, c_void_p,
                                 c_void_p, c_void_p, c_size_t]),
    "mitre_comm_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p), c_void_p]),
    "mitre_comm_connect": (c_int, [c_void_p, c_void_p]),
    "mitre_comm_destroy": (c_int, [c_void_p]),
    "mitre_sharded_chunk_sums": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_size_t]),
    "mitre_sharded_draw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_i64, c_i64, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "mitre_sharded_select": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64,
                                     c_void_p, c_size_t, c_void_p, c_void_p]),
    "mitre_draw_from_tiles": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_double, c_int, c_void_p,
                                      c_void_p, c_void_p, c_size_t]),
}

_lib = None


class MitreError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libmitre_b200.so not found at %s -- build it with `python -m mitre_b200.build`; "
                "mitre_b200 has no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=""):
    if code != MITRE_OK:
        msg = lib().mitre_error_string(int(code)).decode()
        raise MitreError("%s failed: %s (code %d)" % (what or "libmitre_b200 call", msg, code))
