"""ctypes binding of the C ABI in include/sdfa_b200.h (lib/libsdfa_b200.so).

The library is the product; there is no Python or CPU fallback.  If it has not been built
(``make -C sdfa-2019_b200``, or ``__graft_entry__.build()``) importing this module raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDFA_LIB") or os.path.join(_HERE, "..", "lib", "libsdfa_b200.so")   # SDFA_LIB: A/B builds

OK, ERR_ARG, ERR_FACTOR, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4, 5


class SdfaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[sdfa_b200 error {code}] {msg}")
        self.code = code


def _load():
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA library first (make -C sdfa-2019_b200). "
            "This package has no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, ci, cd, cl, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_long, ctypes.c_longlong
    pi = ctypes.POINTER(ctypes.c_int)
    sig = {
        "sdfa_create": ([ctypes.POINTER(vp), vp, ci, vp, ci, vp, ci, vp, cd, ci], ci),
        "sdfa_create_with": ([ctypes.POINTER(vp), vp, ci, vp, ci, vp, ci, vp, cd, ci, ctypes.c_char_p], ci),
        "sdfa_set_option": ([vp, ctypes.c_char_p, cll], ci),
        "sdfa_destroy": ([vp], None),
        "sdfa_info": ([vp, pi, pi, pi, pi, pi, pi, ctypes.POINTER(cll)], ci),
        "sdfa_last_error": ([], ctypes.c_char_p),
        "sdfa_set_constraint_positions": ([vp, vp], ci),
        "sdfa_set_correspondences": ([vp, vp, vp, ci], ci),
        "sdfa_reconstruct_dev": ([vp, vp, cll, ci, vp, vp], ci),
        "sdfa_reconstruct_host": ([vp, vp, ci, vp], ci),
        "sdfa_get_mesh_f64": ([vp, vp, cll, vp, vp, vp, cll, vp], ci),
        "sdfa_get_mesh_from_dm_f64": ([vp, vp, cll, vp, vp], ci),
        "sdfa_set_pca": ([vp, vp, vp, ci, vp, vp, ci], ci),
        "sdfa_decode_reconstruct_dev": ([vp, vp, vp, ci, vp, vp], ci),
        "sdfa_decode_reconstruct_host": ([vp, vp, vp, ci, vp], ci),
        "sdfa_free_vertices": ([vp, vp, ci], ci),
        "sdfa_reconstruct_free_dev": ([vp, vp, cll, ci, vp, vp], ci),
        "sdfa_reconstruct_free_host": ([vp, vp, ci, vp], ci),
        "sdfa_decode_reconstruct_free_dev": ([vp, vp, vp, ci, vp, vp], ci),
        "sdfa_decode_reconstruct_free_host": ([vp, vp, vp, ci, vp], ci),
        "sdfa_expand_free_dev": ([vp, vp, ci, vp, vp], ci),
        "sdfa_decode_dgrad_dev": ([vp, vp, vp, ci, vp, vp], ci),
        "sdfa_decode_compact_dev": ([vp, vp, vp, ci, vp, vp], ci),
        "sdfa_compact_layout": ([vp, vp, ci], ci),
        "sdfa_get_deform_grad_host": ([vp, vp, ci, vp, ci, cd, ci, ci, vp], ci),
        "sdfa_deform_grad_batch_dev": ([vp, vp, ci, vp, ci, ci, cd, ci, vp, vp], ci),
        "sdfa_seek_dev": ([vp, ci, cll, vp, vp, ci, vp, vp], ci),
        "sdfa_launch_count": ([], cll),
        "sdfa_set_timing": ([vp, ci], ci),
        "sdfa_last_timing": ([vp, vp], ci),
        "sdfa_debug_get": ([vp, ctypes.c_char_p, vp, cll], cll),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return lib


lib = _load()
EXPORTS = ["sdfa_create", "sdfa_create_with", "sdfa_set_option", "sdfa_destroy", "sdfa_info", "sdfa_last_error", "sdfa_set_constraint_positions",
           "sdfa_set_correspondences", "sdfa_reconstruct_dev", "sdfa_reconstruct_host", "sdfa_get_mesh_f64",
           "sdfa_get_mesh_from_dm_f64", "sdfa_set_pca", "sdfa_decode_reconstruct_dev",
           "sdfa_decode_reconstruct_host", "sdfa_free_vertices", "sdfa_reconstruct_free_dev",
           "sdfa_reconstruct_free_host", "sdfa_decode_reconstruct_free_dev", "sdfa_decode_reconstruct_free_host",
           "sdfa_expand_free_dev", "sdfa_decode_dgrad_dev", "sdfa_decode_compact_dev", "sdfa_compact_layout",
           "sdfa_get_deform_grad_host", "sdfa_deform_grad_batch_dev", "sdfa_seek_dev",
           "sdfa_launch_count", "sdfa_set_timing", "sdfa_last_timing", "sdfa_debug_get"]


def check(rc):
    if rc != OK:
        raise SdfaError(rc, lib.sdfa_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Host numpy array or integer device pointer -> c_void_p (None stays NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(int(a))


_DEBUG_DTYPES = {
    "perm": np.int32, "parent": np.int32, "free_to_vi": np.int32, "l_colptr": np.int32, "l_rowidx": np.int32,
    "l_val": np.float64, "m_colptr": np.int32, "m_rowidx": np.int32, "m_val": np.float64, "x_base": np.float64,
    "active_eq": np.int32, "tri_u": np.float64, "prog": np.uint8, "stage_off": np.uint32, "io_desc": np.uint32, "io_phase": np.uint32, "eq_src": np.int32,
    "asm_eq_id": np.int32, "asm_eq_u": np.float32, "asm_row_perm": np.int32, "asm_eq_rows": np.int16,
    "asm_colour_ptr": np.int32, "solve_prof": np.int64, "asm_blocks": np.int32, "stats": np.int64,
    "scratch_row": np.int32, "compact_tile": np.int32, "decode_kind": np.int32, "ts_mma": np.uint8, "ts_epi": np.uint8, "ts_matrix": np.uint8, "ts_chunk_off": np.uint32,
    "ts_why_not": np.uint8, "ts_stats": np.int64,
}


def debug_get(handle, what: str) -> np.ndarray:
    n = lib.sdfa_debug_get(handle, what.encode(), None, 0)
    if n < 0:
        raise KeyError(what)
    buf = np.empty(n, dtype=np.uint8)
    lib.sdfa_debug_get(handle, what.encode(), ptr(buf), n)
    return buf.view(_DEBUG_DTYPES[what])
