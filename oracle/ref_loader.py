"""TEST INFRASTRUCTURE ONLY -- loaders for the compiled, unmodified reference in oracle/_ref.

``load_ref_module()``  -> the reference's own pybind module (deformation/cpp/src/pybind.cpp:129-153),
                          loaded WITHOUT registering it in sys.modules so it can coexist with the
                          product's drop-in ``deformation`` package in one process.
``RefSolver``          -> ctypes view of oracle/ref_shim.cpp (one TriangleDeformation per instance;
                          ``get_mesh_batch`` runs one instance per host thread for the CPU baseline).
Both return/raise cleanly when oracle/_ref has not been built (``make -C oracle``).
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def ref_available() -> bool:
    return bool(glob.glob(os.path.join(REF_DIR, "deformation*.so"))) and os.path.exists(
        os.path.join(REF_DIR, "libsdfa_ref.so"))


def load_ref_module():
    paths = glob.glob(os.path.join(REF_DIR, "deformation*.so"))
    if not paths:
        raise FileNotFoundError("oracle/_ref/deformation*.so not built (run `make -C oracle`)")
    spec = importlib.util.spec_from_file_location("deformation", paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_lib = None


def _shim():
    global _lib
    if _lib is None:
        p = os.path.join(REF_DIR, "libsdfa_ref.so")
        if not os.path.exists(p):
            raise FileNotFoundError("oracle/_ref/libsdfa_ref.so not built (run `make -C oracle`)")
        lib = ctypes.CDLL(p)
        vp, ci, cd, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_long
        lib.sdfa_ref_create.restype = vp
        lib.sdfa_ref_destroy.argtypes = [vp]
        lib.sdfa_ref_set_target.argtypes = [vp, vp, ci, vp, ci, vp, ci, vp, cd]
        lib.sdfa_ref_set_target.restype = ci
        lib.sdfa_ref_get_mesh.argtypes = [vp, vp, vp, vp, vp, vp]
        lib.sdfa_ref_get_mesh.restype = ci
        lib.sdfa_ref_get_mesh_from_dm.argtypes = [vp, vp, vp, vp]
        lib.sdfa_ref_get_mesh_from_dm.restype = ci
        lib.sdfa_ref_get_deform_grad.argtypes = [vp, vp, vp, vp, ci, vp, ci, cd]
        lib.sdfa_ref_get_deform_grad.restype = ci
        lib.sdfa_ref_get_deform_mat.argtypes = [vp, vp, vp, vp, ci, vp, ci, cd]
        lib.sdfa_ref_get_deform_mat.restype = ci
        lib.sdfa_ref_get_mesh_batch.argtypes = [vp, ci, vp, cl, cl, vp, vp, cl]
        lib.sdfa_ref_get_mesh_batch.restype = cd
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class RefSolver:
    """n_instances independent reference solvers sharing one template (one per host thread)."""

    def __init__(self, n_instances: int = 1):
        self.lib = _shim()
        self.handles = [self.lib.sdfa_ref_create() for _ in range(n_instances)]
        self.n_verts = self.n_tris = self.n_cnsts = 0

    def close(self):
        for h in self.handles:
            self.lib.sdfa_ref_destroy(h)
        self.handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_target(self, verts, faces, cnsts=(), corrs=(), reg=1e-10) -> bool:
        V = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        F = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        c = np.ascontiguousarray(cnsts, dtype=np.uint32).reshape(-1)
        cc = np.ascontiguousarray(corrs, dtype=np.uint32).reshape(-1)
        ok = True
        for h in self.handles:
            ok &= bool(self.lib.sdfa_ref_set_target(h, _ptr(V), len(V), _ptr(F), len(F), _ptr(c), len(c),
                                                    _ptr(cc) if cc.size else None, float(reg)))
        self.n_verts, self.n_tris, self.n_cnsts = len(V), len(F), len(c)
        return ok

    def get_mesh(self, deform_grad, vert_cnsts=(), corr_count=(), corr_faces=()):
        dg = np.ascontiguousarray(deform_grad, dtype=np.float64).reshape(-1)
        C = np.ascontiguousarray(vert_cnsts, dtype=np.float32).reshape(-1)
        cc = np.ascontiguousarray(corr_count, dtype=np.uint32).reshape(-1)
        cf = np.ascontiguousarray(corr_faces, dtype=np.uint32).reshape(-1)
        out = np.empty((self.n_verts, 3), dtype=np.float32)
        self.lib.sdfa_ref_get_mesh(self.handles[0], _ptr(out), _ptr(dg), _ptr(C) if C.size else None,
                                   _ptr(cc) if cc.size else None, _ptr(cf) if cc.size else None)
        return out

    def get_mesh_batch(self, dgrad_f32, vert_cnsts, n_threads=None):
        """Returns (verts[N,n_verts,3] f32, seconds of the slowest thread)."""
        dg = np.ascontiguousarray(dgrad_f32, dtype=np.float32)
        n = dg.shape[0]
        dg = dg.reshape(n, -1)
        C = np.ascontiguousarray(vert_cnsts, dtype=np.float32).reshape(-1)
        out = np.empty((n, self.n_verts, 3), dtype=np.float32)
        nt = len(self.handles) if n_threads is None else min(n_threads, len(self.handles))
        arr = (ctypes.c_void_p * nt)(*self.handles[:nt])
        secs = self.lib.sdfa_ref_get_mesh_batch(arr, nt, _ptr(dg), n, dg.shape[1],
                                                _ptr(C) if C.size else None, _ptr(out), self.n_verts * 3)
        return out, float(secs)

    def get_deform_grad(self, verts_a, verts_b, faces, eps=1e-6):
        A = np.ascontiguousarray(verts_a, dtype=np.float32).reshape(-1, 3)
        B = np.ascontiguousarray(verts_b, dtype=np.float32).reshape(-1, 3)
        F = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        out = np.empty(len(F) * 9, dtype=np.float64)
        self.lib.sdfa_ref_get_deform_grad(self.handles[0], _ptr(out), _ptr(A), _ptr(B), len(A), _ptr(F),
                                          len(F), float(eps))
        return out
