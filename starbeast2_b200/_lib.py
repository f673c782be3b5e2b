"""ctypes binding of libsb2gpu.so (include/sb2gpu.h).  No torch types cross this boundary.

The library is built in-tree (starbeast2_b200/lib/libsb2gpu.so) by starbeast2_b200.build.  There is no
fallback: if the library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB2_LIB") or os.path.join(HERE, "lib", "libsb2gpu.so")   # SB2_LIB: A/B against another build
HEADER = os.path.join(os.path.dirname(HERE), "include", "sb2gpu.h")

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_vp = C.c_void_p


class Sb2Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("exact_order", C.c_int32), ("profile", C.c_int32), ("reserved", C.c_int32 * 5)]


class Sb2Error(RuntimeError):
    pass


_lib = None


def header_symbols() -> list:
    """Every SB2_API function the header declares."""
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SB2_API\s+[\w\s\*]+?\b(sb2_\w+)\s*\(", src)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sb2Error(f"{LIB_PATH} is missing: run `python -m starbeast2_b200.build` (needs nvcc); "
                       "there is no CPU fallback for the likelihood engine")
    L = C.CDLL(LIB_PATH)
    E = _vp
    sig = {
        "sb2_create": (C.c_int, [C.POINTER(Sb2Config), C.POINTER(_vp)]),
        "sb2_destroy": (None, [E]),
        "sb2_last_error": (C.c_char_p, [E]),
        "sb2_add_locus": (C.c_int, [E, C.c_int32, C.c_int32, _u8p, _i32p, C.c_int32, _i32p, C.c_double]),
        "sb2_add_locus_k": (C.c_int, [E, C.c_int32, C.c_int32, C.c_int32, _u8p, _i32p, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_void_p]),
        "sb2_finalize": (C.c_int, [E, C.c_int32, C.c_int32]),
        "sb2_set_species_tree": (C.c_int, [E, _i32p, _i32p, _i32p, _f64p]),
        "sb2_set_gene_tree": (C.c_int, [E, C.c_int32, _i32p, _i32p, _i32p, _f64p]),
        "sb2_set_gene_trees": (C.c_int, [E, C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _f64p]),
        "sb2_set_subst_model": (C.c_int, [E, C.c_int32, C.c_int32, _f64p]),
        "sb2_set_site_model": (C.c_int, [E, C.c_int32, _f64p, _f64p, C.c_double]),
        "sb2_set_clock": (C.c_int, [E, C.c_int32, C.c_int32, C.c_double, _vp]),
        "sb2_set_species_rates": (C.c_int, [E, _f64p]),
        "sb2_set_pop_model": (C.c_int, [E, C.c_int32, _vp, _vp, C.c_double, C.c_double]),
        "sb2_eval_coalescent": (C.c_int, [E, _vp, _vp, _vp]),
        "sb2_eval_likelihood": (C.c_int, [E, _vp, _vp]),
        "sb2_eval_posterior": (C.c_int, [E, _vp, _vp, _vp, _vp]),
        "sb2_eval_begin": (C.c_int, [E]),
        "sb2_eval_end": (C.c_int, [E, _vp, _vp, _vp, _vp]),
        "sb2_reduce_vector": (C.c_int, [E, C.POINTER(_vp), C.POINTER(C.c_int32)]),
        "sb2_bind_reduce_vector": (C.c_int, [E, _vp]),
        "sb2_read_reduce_vector": (C.c_int, [E, _f64p]),
        "sb2_write_reduce_vector": (C.c_int, [E, _f64p]),
        "sb2_comm_handle": (C.c_int, [E, _vp]),
        "sb2_comm_connect": (C.c_int, [E, C.c_int32, C.c_int32, _vp]),
        "sb2_set_stream": (C.c_int, [E, _vp]),
        "sb2_store": (C.c_int, [E]),
        "sb2_restore": (C.c_int, [E]),
        "sb2_accept": (C.c_int, [E]),
        "sb2_mark_all_dirty": (C.c_int, [E]),
        "sb2_get_branch_rates": (C.c_int, [E, C.c_int32, _f64p]),
        "sb2_get_embedding": (C.c_int, [E, C.c_int32, _vp, _vp, _vp, _vp]),
        "sb2_get_partials": (C.c_int, [E, C.c_int32, C.c_int32, _vp, _vp]),
        "sb2_get_matrices": (C.c_int, [E, C.c_int32, C.c_int32, _f64p]),
        "sb2_get_pattern_logl": (C.c_int, [E, C.c_int32, _f64p]),
        "sb2_connecting_nodes": (C.c_int, [E, C.c_int32, _vp, _f64p, C.POINTER(C.c_int64)]),
        "sb2_max_height": (C.c_int, [E, _u8p, C.POINTER(C.c_double)]),
        "sb2_compress_alignment": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _u8p, _u8p, _i32p, _vp, C.POINTER(C.c_int32)]),
        "sb2_compress_last_error": (C.c_char_p, []),
        "sb2_num_loci": (C.c_int32, [E]),
        "sb2_launch_count": (C.c_int64, [E]),
        "sb2_last_node_updates": (C.c_int64, [E]),
        "sb2_kernel_time": (C.c_int, [E, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "sb2_set_profile": (C.c_int, [E, C.c_int32]),
        "sb2_device_bytes": (C.c_int64, [E]),
        "sb2_step_phase_times": (C.c_int, [E, C.POINTER(C.c_double), C.c_int32]),
        "sb2_exchange_wait_us": (C.c_double, [E]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._sb2_signatures = sig
    _lib = L
    return L


_raw = None


def load_raw() -> C.CDLL:
    """A second handle on the same library whose hot entry points take raw addresses (c_void_p) instead of checked numpy
    arrays: what a JNI / FFM caller does -- resolve the buffers once, then just call.  Used by bench.py's e2e loop so that
    the number reflects the C ABI, not numpy argument validation."""
    global _raw
    if _raw is not None:
        return _raw
    load()
    L = C.CDLL(LIB_PATH)
    E = _vp
    for name, args in {
        "sb2_set_species_tree": [E, _vp, _vp, _vp, _vp],
        "sb2_set_gene_trees": [E, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp],
        "sb2_set_species_rates": [E, _vp],
        "sb2_set_pop_model": [E, C.c_int32, _vp, _vp, C.c_double, C.c_double],
        "sb2_mark_all_dirty": [E],
        "sb2_eval_posterior": [E, _vp, _vp, _vp, _vp],
        "sb2_eval_begin": [E],
        "sb2_eval_end": [E, _vp, _vp, _vp, _vp],
        "sb2_eval_coalescent": [E, _vp, _vp, _vp],
        "sb2_eval_likelihood": [E, _vp, _vp],
        "sb2_store": [E],
        "sb2_restore": [E],
        "sb2_accept": [E],
    }.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _raw = L
    return L
