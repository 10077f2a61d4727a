"""Drop-in ``deformation`` package backed by the B200-native CUDA library (no CPU fallback).

Same names, keyword arguments, defaults, dtypes and return shapes as the reference's pybind module
(reference: deformation/cpp/src/pybind.cpp:129-153, imported through deformation/__init__.py:6-13):

    set_target(verts, faces, cnsts=[], corrs=[], reg=1e-10) -> bool
    is_same(num_verts, num_faces, num_cnsts) -> bool
    get_mesh(deform_grad, vert_cnsts=[], corr_count=[], corr_faces=[]) -> float32[n_verts,3]
    get_mesh_from_dg            (alias of get_mesh)
    get_mesh_from_dm(deform_mat, vert_cnsts=[]) -> float32[n_verts,3]
    get_deform_grad(verts_a, verts_b, faces, eps=1e-6) -> float64[9*n_tris]
    get_deform_mat(verts_a, verts_b, faces, eps=1e-6)  -> float64[9*n_tris]

so ``speech_anime/viewer/frame.py:42,118-137`` runs unchanged against it.  Like the reference these act on
one process-global solver (pybind.cpp:10).  Differences, all deliberate:
  * argument errors raise ``SdfaError`` instead of logging and calling exit(1) (log.hpp:32-33);
  * arithmetic is float32 on the GPU around an fp64-factorised system: results match the reference to
    <= 1e-6 x bbox diagonal, not bit for bit.
New, batched entry points (what the hot loops of speech_anime/model/model.py:201-212 and
viewer/video.py:220-277 should call instead of one frame at a time):

    get_mesh_batch(deform_grads[N, 9*n_tris]) -> float32[N, n_verts, 3]      (numpy or torch.cuda tensors)
    set_pca(compT_scale, means_scale, compT_rotat, means_rotat)
    decode_and_get_mesh(coeff_scale[N,Ks], coeff_rotat[N,Kr]) -> float32[N, n_verts, 3]
    Reconstructor(...)            explicit-handle version of all of the above (one per device/template)
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _native
from ._native import SdfaError, check, lib, ptr

__all__ = ["set_target", "is_same", "get_mesh", "get_mesh_from_dg", "get_mesh_from_dm", "get_deform_grad",
           "get_deform_mat", "get_deform_grad_batch", "get_mesh_batch", "set_pca", "decode_and_get_mesh", "Reconstructor", "SdfaError"]


def _f32c(a, what):
    """py::array_t<float, c_style> semantics: C-contiguous + forcecast (pybind.cpp:14)."""
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32c(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _default_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class Reconstructor:
    """One template (+ constraint set, correspondences, PCA basis) on one CUDA device.

    Construction = TriangleDeformation::setStaticTarget (deform_triangle_impl.hpp:7-142): builds
    A^T A + reg*I, factors it (fp64, host), schedules the sweeps and uploads the plan.
    ``device=-1`` keeps everything on the host for plan inspection; compute calls then raise.
    """

    def __init__(self, verts, faces, cnsts=(), corrs=(), reg=1e-10, device=None, solver=None, options=None):
        """``solver``: None = tensor-core solve when the template fits it, else the SIMT sweeps;
        "simt" / "tensor" force one.  ``options``: further ``sdfa_create_with`` options as a dict
        (include/sdfa_b200.h), e.g. ``{"pipe_chunk": 4096}``."""
        V = _f32c(verts, "verts")
        F = _u32c(faces)
        c = _u32c(cnsts).reshape(-1)
        cc = _u32c(corrs).reshape(-1)
        if not (1 <= V.ndim <= 2 and 1 <= F.ndim <= 2):                      # pybind.cpp:20-21
            raise SdfaError(_native.ERR_ARG, "verts/faces must be 1-D or 2-D")
        V = V.reshape(-1, 3)
        F = F.reshape(-1, 3)
        if cc.size not in (0, len(F)):
            raise SdfaError(_native.ERR_ARG, "corrs must be empty or have one count per face")
        self.n_verts, self.n_tris, self.n_cnsts = len(V), len(F), len(c)
        self.device = _default_device() if device is None else int(device)
        self._h = ctypes.c_void_p()
        self._verts, self._faces, self._cnsts = V, F, c
        opts = dict(options or {})
        if solver is not None:
            opts["solver"] = solver
        text = ";".join(f"{k}={v}" for k, v in opts.items())
        check(lib.sdfa_create_with(ctypes.byref(self._h), ptr(V), len(V), ptr(F), len(F), ptr(c) if c.size else None,
                                   len(c), ptr(cc) if cc.size else None, float(reg), self.device,
                                   text.encode() if text else None))
        nf, ne, na, nnz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
        check(lib.sdfa_info(self._h, None, None, None, ctypes.byref(nf), ctypes.byref(ne), ctypes.byref(na),
                            ctypes.byref(nnz)))
        self.n_free, self.n_eq, self.n_active, self.nnz_l = nf.value, ne.value, na.value, nnz.value
        self.n_src_tris = self.n_tris
        self._corr_key = None
        self._has_pca = False

    # -------------------------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.sdfa_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def debug(self, what):
        return _native.debug_get(self._h, what)

    def set_option(self, name, value):
        """Run-time options of the handle (``sdfa_set_option``): ``pipe_chunk``."""
        check(lib.sdfa_set_option(self._h, name.encode(), int(value)))

    def _torch_out(self, out, n, like, free_only=False):
        """Validated (or freshly allocated) [n, rows, 3] float32 output tensor on the handle's device
        (rows = n_verts, or n_free with ``free_only``)."""
        import torch
        rows = self.n_free if free_only else self.n_verts
        if like.device.index != self.device:
            raise SdfaError(_native.ERR_ARG, f"input is on {like.device}, the handle on cuda:{self.device}")
        if out is None:
            return torch.empty((n, rows, 3), dtype=torch.float32, device=like.device)
        if (not _is_torch(out) or out.dtype != torch.float32 or out.device != like.device or not out.is_contiguous()
                or out.numel() != n * rows * 3):
            raise SdfaError(_native.ERR_ARG, f"out must be a contiguous float32 tensor of {n}x{rows}x3 on {like.device}")
        return out

    def _numpy_out(self, out, n, free_only=False):
        rows = self.n_free if free_only else self.n_verts
        if out is None:
            return np.empty((n, rows, 3), dtype=np.float32)
        if (not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags.c_contiguous
                or out.size != n * rows * 3):
            raise SdfaError(_native.ERR_ARG, f"out must be a C-contiguous float32 array of {n}x{rows}x3")
        return out

    @property
    def free_vertices(self):
        """Vertex index of every row of a ``free_only`` result (ascending; the vertices that are not constrained)."""
        ids = np.empty(self.n_free, dtype=np.int32)
        lib.sdfa_free_vertices(self._h, ptr(ids), self.n_free)
        return ids

    def expand_free(self, free_rows, out=None, stream=None):
        """[N, n_free, 3] free rows (torch.cuda) -> [N, n_verts, 3] with the current constraint positions filled in."""
        import torch
        x = free_rows.contiguous()
        if x.dtype != torch.float32 or not x.is_cuda or x.numel() % (self.n_free * 3):
            raise SdfaError(_native.ERR_ARG, "expand_free: float32 CUDA tensor [N, n_free, 3]")
        n = x.numel() // (self.n_free * 3)
        out = self._torch_out(out, n, x)
        s = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        check(lib.sdfa_expand_free_dev(self._h, ptr(x.data_ptr()), n, ptr(out.data_ptr()), ptr(s)))
        return out

    # ---------------------------------------------------------------------------------- state
    def set_constraint_positions(self, vert_cnsts=None):
        if vert_cnsts is None or np.size(vert_cnsts) == 0:
            check(lib.sdfa_set_constraint_positions(self._h, None))
            return
        C = _f32c(vert_cnsts, "vert_cnsts").reshape(-1, 3)
        if len(C) != self.n_cnsts:
            raise SdfaError(_native.ERR_ARG, f"vert_cnsts has {len(C)} rows, template has {self.n_cnsts} constraints")
        check(lib.sdfa_set_constraint_positions(self._h, ptr(C)))

    def set_correspondences(self, corr_count=(), corr_faces=(), n_src_tris=None):
        cc, cf = _u32c(corr_count).reshape(-1), _u32c(corr_faces).reshape(-1)
        if cc.size == 0:
            check(lib.sdfa_set_correspondences(self._h, None, None, self.n_tris))
            self.n_src_tris = self.n_tris
            return
        if n_src_tris is None:
            n_src_tris = int(cf.max()) + 1 if cf.size else 0
        if cc.size != self.n_tris or cf.size < self.n_eq:
            raise SdfaError(_native.ERR_ARG, "corr_count/corr_faces sizes do not match the template")
        check(lib.sdfa_set_correspondences(self._h, ptr(cc), ptr(cf), int(n_src_tris)))
        self.n_src_tris = int(n_src_tris)

    def set_pca(self, compT_scale, means_scale, compT_rotat, means_rotat):
        """PcaInversion buffers (output_module.py:94-113): compT [out, K], means [out]."""
        cs, ms = _f32c(compT_scale, "compT_scale"), _f32c(means_scale, "means_scale").reshape(-1)
        cr, mr = _f32c(compT_rotat, "compT_rotat"), _f32c(means_rotat, "means_rotat").reshape(-1)
        nt = self.n_src_tris
        if cs.shape[0] != nt * 6 or cr.shape[0] != nt * 3 or ms.size != nt * 6 or mr.size != nt * 3:
            raise SdfaError(_native.ERR_ARG, "PCA basis shapes must be [n_tris*6,Ks], [n_tris*6], [n_tris*3,Kr], [n_tris*3]")
        self.k_scale, self.k_rotat = cs.shape[1], cr.shape[1]
        check(lib.sdfa_set_pca(self._h, ptr(cs), ptr(ms), self.k_scale, ptr(cr), ptr(mr), self.k_rotat))
        self._has_pca = True

    # ------------------------------------------------------------------------ single frame (legacy)
    def get_mesh(self, deform_grad, vert_cnsts=(), corr_count=(), corr_faces=()):
        dg = np.ascontiguousarray(deform_grad, dtype=np.float64).reshape(-1)
        C = _f32c(vert_cnsts, "vert_cnsts").reshape(-1)
        cc, cf = _u32c(corr_count).reshape(-1), _u32c(corr_faces).reshape(-1)
        out = np.empty((self.n_verts, 3), dtype=np.float32)
        if C.size not in (0, self.n_cnsts * 3):
            raise SdfaError(_native.ERR_ARG, "vert_cnsts must have one row per constraint")
        check(lib.sdfa_get_mesh_f64(self._h, ptr(dg), dg.size, ptr(C) if C.size else None,
                                    ptr(cc) if cc.size else None, ptr(cf) if cc.size else None, cf.size, ptr(out)))
        self.n_src_tris = dg.size // 9
        return out

    def get_mesh_from_dm(self, deform_mat, vert_cnsts=()):
        dm = np.ascontiguousarray(deform_mat, dtype=np.float64).reshape(-1)
        C = _f32c(vert_cnsts, "vert_cnsts").reshape(-1)
        out = np.empty((self.n_verts, 3), dtype=np.float32)
        check(lib.sdfa_get_mesh_from_dm_f64(self._h, ptr(dm), dm.size, ptr(C) if C.size else None, ptr(out)))
        return out

    # ----------------------------------------------------------------------------------- batched
    def get_mesh_batch(self, deform_grads, out=None, stream=None, free_only=False):
        """[N, 9*n_src_tris] float32 -> [N, n_verts, 3] float32.  torch.cuda tensors stay on the device
        (stream-ordered, no synchronisation); numpy arrays go through pinned-size staging copies.
        ``free_only``: return only the free vertices, [N, n_free, 3] (rows = ``free_vertices``): the constrained
        rows are the constants given as ``vert_cnsts``, so a host copy or gather moves a quarter of the bytes."""
        if _is_torch(deform_grads):
            import torch
            x = deform_grads
            if x.dtype != torch.float32 or not x.is_cuda:
                raise SdfaError(_native.ERR_ARG, "torch input must be a float32 CUDA tensor")
            x = x.reshape(x.shape[0], -1)
            if x.stride(1) != 1:
                x = x.contiguous()
            if x.shape[1] != self.n_src_tris * 9:
                raise SdfaError(_native.ERR_ARG, f"expected {self.n_src_tris * 9} values per frame, got {x.shape[1]}")
            n = x.shape[0]
            out = self._torch_out(out, n, x, free_only)
            s = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
            fn = lib.sdfa_reconstruct_free_dev if free_only else lib.sdfa_reconstruct_dev
            check(fn(self._h, ptr(x.data_ptr()), x.stride(0), n, ptr(out.data_ptr()), ptr(s)))
            return out
        x = _f32c(deform_grads, "deform_grads")
        x = x.reshape(x.shape[0], self.n_src_tris * 9) if (x.ndim > 1 and x.size == 0) else (
            x.reshape(x.shape[0], -1) if x.ndim > 1 else x.reshape(1, -1))
        if x.shape[1] != self.n_src_tris * 9:
            raise SdfaError(_native.ERR_ARG, f"expected {self.n_src_tris * 9} values per frame, got {x.shape[1]}")
        res = self._numpy_out(out, x.shape[0], free_only)
        fn = lib.sdfa_reconstruct_free_host if free_only else lib.sdfa_reconstruct_host
        check(fn(self._h, ptr(x), x.shape[0], ptr(res)))
        return res

    def decode_and_get_mesh(self, coeff_scale, coeff_rotat, out=None, stream=None, free_only=False):
        """PCA coefficients -> vertices: F.linear x2 + interleave + reconstruction in one call
        (``free_only`` as in ``get_mesh_batch``)."""
        if not self._has_pca:
            raise SdfaError(_native.ERR_STATE, "set_pca() first")
        if _is_torch(coeff_scale):
            import torch
            a, b = coeff_scale.contiguous(), coeff_rotat.contiguous()
            if a.dtype != torch.float32 or b.dtype != torch.float32 or not a.is_cuda or not b.is_cuda:
                raise SdfaError(_native.ERR_ARG, "torch inputs must be float32 CUDA tensors")
            a, b = a.reshape(-1, self.k_scale), b.reshape(-1, self.k_rotat)
            n = a.shape[0]
            if b.shape[0] != n or b.device != a.device:
                raise SdfaError(_native.ERR_ARG, "coefficient batches differ in length or device")
            out = self._torch_out(out, n, a, free_only)
            s = torch.cuda.current_stream(a.device).cuda_stream if stream is None else stream
            fn = lib.sdfa_decode_reconstruct_free_dev if free_only else lib.sdfa_decode_reconstruct_dev
            check(fn(self._h, ptr(a.data_ptr()), ptr(b.data_ptr()), n, ptr(out.data_ptr()), ptr(s)))
            return out
        a = _f32c(coeff_scale, "coeff_scale").reshape(-1, self.k_scale)
        b = _f32c(coeff_rotat, "coeff_rotat").reshape(-1, self.k_rotat)
        if len(a) != len(b):
            raise SdfaError(_native.ERR_ARG, "coefficient batches differ in length")
        res = self._numpy_out(out, len(a), free_only)
        fn = lib.sdfa_decode_reconstruct_free_host if free_only else lib.sdfa_decode_reconstruct_host
        check(fn(self._h, ptr(a), ptr(b), len(a), ptr(res)))
        return res

    def decode_dgrad(self, coeff_scale, coeff_rotat, stream=None):
        """data_to_anime_feat (model.py:246-257): coefficients -> full-layout dgrad [N, 9*n_tris] (torch.cuda only)."""
        import torch
        a = coeff_scale.contiguous().reshape(-1, self.k_scale)
        b = coeff_rotat.contiguous().reshape(-1, self.k_rotat)
        if a.shape[0] != b.shape[0] or a.dtype != torch.float32 or b.dtype != torch.float32 or not (a.is_cuda and b.is_cuda):
            raise SdfaError(_native.ERR_ARG, "decode_dgrad: float32 CUDA coefficient tensors of equal length")
        out = torch.empty((a.shape[0], self.n_src_tris * 9), dtype=torch.float32, device=a.device)
        s = torch.cuda.current_stream(a.device).cuda_stream if stream is None else stream
        check(lib.sdfa_decode_dgrad_dev(self._h, ptr(a.data_ptr()), ptr(b.data_ptr()), a.shape[0],
                                        ptr(out.data_ptr()), ptr(s)))
        return out

    def compact_layout(self):
        """Map of the internal compact dgrad: slot i = source_triangle*9 + component, -1 = padding."""
        n = lib.sdfa_compact_layout(self._h, None, 0)
        out = np.empty(n, dtype=np.int32)
        lib.sdfa_compact_layout(self._h, ptr(out), n)
        return out

    def decode_compact(self, coeff_scale, coeff_rotat, stream=None):
        """Coefficients -> compact dgrad (torch.cuda; tcgen05 kernel).  The kernel writes the frame-tiled layout
        [ceil(N/T), slots, T] the assembly kernel reads (T = 64 frames per tile); returned here as [N, slots]
        (slots as in compact_layout())."""
        import torch
        a = coeff_scale.contiguous().reshape(-1, self.k_scale)
        b = coeff_rotat.contiguous().reshape(-1, self.k_rotat)
        slots = lib.sdfa_compact_layout(self._h, None, 0)
        n = a.shape[0]
        T = int(self.debug("compact_tile")[0])
        out = torch.zeros(((n + T - 1) // T, slots, T), dtype=torch.float32, device=a.device)
        s = torch.cuda.current_stream(a.device).cuda_stream if stream is None else stream
        check(lib.sdfa_decode_compact_dev(self._h, ptr(a.data_ptr()), ptr(b.data_ptr()), n, ptr(out.data_ptr()), ptr(s)))
        return out.permute(0, 2, 1).reshape(-1, slots)[:n]

    # ------------------------------------------------------------------------------ measurement
    def set_timing(self, enable=True):
        check(lib.sdfa_set_timing(self._h, 1 if enable else 0))

    def last_timing(self):
        ms = (ctypes.c_float * 4)()
        check(lib.sdfa_last_timing(self._h, ms))
        return dict(decode_ms=ms[0], assembly_ms=ms[1], solve_ms=ms[2], output_ms=ms[3])


# =================================================================================================
# The process-global singleton interface of the reference (pybind.cpp:10: gDeformManager).
_global: Reconstructor | None = None
_counts = (0, 0, 0)


def set_target(verts, faces, cnsts=(), corrs=(), reg=1e-10) -> bool:
    """SetTarget (pybind.cpp:13-33).  Returns False when the factorisation fails (impl.hpp:134-139)."""
    global _global, _counts
    if _global is not None:
        _global.close()
        _global = None
    try:
        _global = Reconstructor(verts, faces, cnsts, corrs, reg)
    except SdfaError as e:
        if e.code == _native.ERR_FACTOR:
            print(f"solver error: {e}")
            return False
        raise
    _counts = (_global.n_verts, _global.n_tris, _global.n_cnsts)
    return True


def is_same(num_verts, num_faces, num_cnsts) -> bool:
    """IsSame (pybind.cpp:119-126): compares the three counts only."""
    return _counts == (int(num_verts), int(num_faces), int(num_cnsts))


def _need_target():
    if _global is None:
        raise SdfaError(_native.ERR_STATE, "set_target() has not been called")
    return _global


def get_mesh(deform_grad, vert_cnsts=(), corr_count=(), corr_faces=()):
    """GetMeshFromGrad (pybind.cpp:101-117)."""
    return _need_target().get_mesh(deform_grad, vert_cnsts, corr_count, corr_faces)


get_mesh_from_dg = get_mesh


def get_mesh_from_dm(deform_mat, vert_cnsts=()):
    """GetMeshFromMat (pybind.cpp:60-74)."""
    return _need_target().get_mesh_from_dm(deform_mat, vert_cnsts)


def _inverse(verts_a, verts_b, faces, eps, as_matrix):
    A, B = _f32c(verts_a, "verts_a"), _f32c(verts_b, "verts_b")
    F = _u32c(faces)
    if A.size != B.size:                                                     # pybind.cpp:47,88
        raise SdfaError(_native.ERR_ARG, "verts_a and verts_b differ in size")
    A, B, F = A.reshape(-1, 3), B.reshape(-1, 3), F.reshape(-1, 3)
    out = np.empty(len(F) * 9, dtype=np.float64)
    check(lib.sdfa_get_deform_grad_host(ptr(A), ptr(B), len(A), ptr(F), len(F), float(eps), as_matrix,
                                        _default_device(), ptr(out)))
    return out


def get_deform_grad(verts_a, verts_b, faces, eps=1e-6):
    """GetDeformGrad (pybind.cpp:78-99)."""
    return _inverse(verts_a, verts_b, faces, eps, 0)


def get_deform_mat(verts_a, verts_b, faces, eps=1e-6):
    """GetDeformMat (pybind.cpp:37-58)."""
    return _inverse(verts_a, verts_b, faces, eps, 1)


def get_deform_grad_batch(verts_a, verts_b, faces, eps=1e-6, as_matrix=False):
    """Many meshes against one template on the GPU: verts_a [n_verts,3], verts_b [N,n_verts,3] float32
    (numpy or torch.cuda) -> float32 [N, 9*n_tris] (what generate_dgrad stores, preload.py:765-835)."""
    import torch
    dev = verts_b.device if _is_torch(verts_b) and verts_b.is_cuda else torch.device("cuda", _default_device())
    A = torch.as_tensor(np.ascontiguousarray(verts_a, dtype=np.float32) if not _is_torch(verts_a) else verts_a,
                        dtype=torch.float32, device=dev).reshape(-1, 3).contiguous()
    B = torch.as_tensor(np.ascontiguousarray(verts_b, dtype=np.float32) if not _is_torch(verts_b) else verts_b,
                        dtype=torch.float32, device=dev).reshape(-1, A.shape[0], 3).contiguous()
    Fh = _u32c(faces.cpu().numpy() if _is_torch(faces) else faces).reshape(-1, 3)
    if Fh.size and int(Fh.max()) >= A.shape[0]:
        raise SdfaError(_native.ERR_ARG, "face index out of range")
    Fd = torch.from_numpy(Fh.astype(np.int32)).to(dev)          # same bits as uint32
    out = torch.empty((B.shape[0], len(Fh) * 9), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    check(lib.sdfa_deform_grad_batch_dev(ptr(A.data_ptr()), ptr(B.data_ptr()), A.shape[0], ptr(Fd.data_ptr()), len(Fh),
                                         B.shape[0], float(eps), 1 if as_matrix else 0, ptr(out.data_ptr()), ptr(s)))
    return out


def seek_batch(query_ts, timestamps, sequence, out=None, stream=None):
    """Batched ``saber.stream.seek`` (saber/data/stream/stream.py:20-46) on the device: ``sequence`` is a
    torch.cuda float32 tensor [n_src, ...] sampled at ``timestamps`` (ascending); returns [len(query_ts), ...] with
    every row blended like the reference does for one timestamp.  Seeking PCA coefficients and then calling
    ``decode_and_get_mesh`` equals decoding and seeking the dgrad (the decode is affine)."""
    import torch
    if not _is_torch(sequence) or not sequence.is_cuda:
        raise SdfaError(_native.ERR_ARG, "seek_batch: sequence must be a torch.cuda tensor")
    seq = sequence.contiguous().to(torch.float32)
    t = np.ascontiguousarray(timestamps, dtype=np.float64).reshape(-1)
    q = np.ascontiguousarray(query_ts, dtype=np.float64).reshape(-1)
    if len(t) != seq.shape[0]:
        raise SdfaError(_native.ERR_ARG, "seek_batch: one timestamp per row of the sequence")   # stream.py:22
    width = int(np.prod(seq.shape[1:])) if seq.dim() > 1 else 1
    if out is None:
        out = torch.empty((len(q),) + tuple(seq.shape[1:]), dtype=torch.float32, device=seq.device)
    s = torch.cuda.current_stream(seq.device).cuda_stream if stream is None else stream
    with torch.cuda.device(seq.device):
        check(lib.sdfa_seek_dev(ptr(seq.data_ptr()), seq.shape[0], width, ptr(t), ptr(q), len(q), ptr(out.data_ptr()), ptr(s)))
    return out


def get_mesh_batch(deform_grads, vert_cnsts=None, out=None):
    r = _need_target()
    r.set_constraint_positions(vert_cnsts)
    return r.get_mesh_batch(deform_grads, out=out)


def set_pca(compT_scale, means_scale, compT_rotat, means_rotat):
    _need_target().set_pca(compT_scale, means_scale, compT_rotat, means_rotat)


def decode_and_get_mesh(coeff_scale, coeff_rotat, vert_cnsts=None, out=None):
    r = _need_target()
    r.set_constraint_positions(vert_cnsts)
    return r.decode_and_get_mesh(coeff_scale, coeff_rotat, out=out)
