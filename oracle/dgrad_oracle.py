"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dgrad <-> mesh path.

This module is the *checker* for the CUDA product under ``sdfa-2019_b200/``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
it; the product never does (it has no CPU fallback).

What it restates (numpy float64 + scipy SuperLU, vectorised over triangles), each
function citing the reference code it follows (paths relative to /root/reference):

* ``set_target``       deformation/cpp/src/deform_triangle_impl.hpp:7-142 (setStaticTarget)
                       and :479-511 (_qrFactorize); marshalling per pybind.cpp:13-33
* ``get_mesh``         deform_triangle_impl.hpp:215-310 (getMeshFromDeformationGradients),
                       rotation/utils_rotation.cpp:20-51 (exp); pybind.cpp:101-117
* ``get_mesh_from_dm`` deform_triangle_impl.hpp:382-440
* ``get_deform_grad``  deform_triangle_impl.hpp:144-213, :443-470 (_getTransform,
                       _getGradFromMat), rotation/utils_rotation.cpp:71-175 (log)
* ``get_deform_mat``   deform_triangle_impl.hpp:313-380
* ``is_same``          pybind.cpp:119-126
* ``pca_decode``       speech_anime/modules/output_module.py:115-116 (F.linear) and
                       speech_anime/model/model.py:246-257 (scale(6)+rotation(3) interleave)
* ``seek``             saber/data/stream/stream.py:20-46 (time-linear interpolation)

The reference's solver is Eigen::SparseLU (deform_triangle.hpp:27) whose source is
vendored under deformation/cpp/ext/eigen3; it is *replaced* here by scipy's SuperLU on
the same matrix ``A^T A + reg*I`` -- both are fp64 direct solves of one linear system, so
results agree to fp64 rounding and (after the float32 output cast) to <= 1 ulp(float32).

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4).  This restatement is pinned against the reference itself, compiled
from its own sources by ``oracle/Makefile`` into ``oracle/_ref`` (tests/test_oracle.py),
and against fixtures that build produced (tests/golden/, made by
tests/golden/make_fixtures.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_QR_EPS = 1e-6          # deform_triangle_impl.hpp:482
_LOGEXP_TOL = 1.0e-6    # rotation/utils_rotation.h:8


def _as_f32_verts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert 1 <= a.ndim <= 2                       # pybind.cpp:20
    return a.reshape(-1, 3)


def _as_u32_tris(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    assert 1 <= a.ndim <= 2                       # pybind.cpp:21
    return a.reshape(-1, 3)


def triangle_frames(verts_f32: np.ndarray, tris: np.ndarray):
    """Per-triangle 2x3 matrix U = R^-1 Q^T (deform_triangle_impl.hpp:92-100, :479-511).

    Edges are subtracted in float32 (Eigen::Vector3f v2-v1, :92-97) and then widened.
    Returns U with shape (m, 2, 3) in float64.
    """
    v1 = verts_f32[tris[:, 0]]
    v2 = verts_f32[tris[:, 1]]
    v3 = verts_f32[tris[:, 2]]
    e1 = (v2 - v1).astype(np.float64)             # float32 subtraction, then widen
    e2 = (v3 - v1).astype(np.float64)
    m = len(tris)
    # classical Gram-Schmidt, column 0
    r00 = np.sqrt(np.einsum("ij,ij->i", e1, e1))
    bad0 = r00 < _QR_EPS
    q0 = np.where(bad0[:, None], 0.0, e1 / np.where(bad0, 1.0, r00)[:, None])
    r00 = np.where(bad0, 1.0, r00)
    # column 1
    r01 = np.einsum("ij,ij->i", q0, e2)
    v = e2 - r01[:, None] * q0
    r11 = np.sqrt(np.einsum("ij,ij->i", v, v))
    bad1 = r11 < _QR_EPS
    q1 = np.where(bad1[:, None], 0.0, v / np.where(bad1, 1.0, r11)[:, None])
    r11 = np.where(bad1, 1.0, r11)
    # U = R^-1 Q^T with R = [[r00, r01], [0, r11]]
    U = np.empty((m, 2, 3))
    U[:, 1, :] = q1 / r11[:, None]
    U[:, 0, :] = q0 / r00[:, None] - (r01 / (r00 * r11))[:, None] * q1
    return U


def rotation_exp(logr: np.ndarray) -> np.ndarray:
    """rotation_log_exp::exp(Matrix3d) (utils_rotation.cpp:32-51 -> :20-30), batched.

    logr: (m,3,3) skew matrices.  angle < 1e-6 => identity.  (The skew checks at
    :34-39 / :22-27 never trigger for matrices built at deform_triangle_impl.hpp:232-235.)
    """
    w = np.stack([logr[:, 2, 1], logr[:, 0, 2], logr[:, 1, 0]], axis=1)
    ang = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2])
    small = ang < _LOGEXP_TOL
    K = logr / np.where(small, 1.0, ang)[:, None, None]
    eye = np.eye(3)[None]
    R = eye + np.sin(ang)[:, None, None] * K + (1.0 - np.cos(ang))[:, None, None] * (K @ K)
    R[small] = np.eye(3)
    return R


def dgrad_to_transforms(dg: np.ndarray) -> np.ndarray:
    """T_i = exp(logR_i) * S_i for dg of shape (m,9) float64 (deform_triangle_impl.hpp:226-244)."""
    m = dg.shape[0]
    logr = np.zeros((m, 3, 3))
    logr[:, 0, 1] = dg[:, 6]; logr[:, 0, 2] = dg[:, 7]; logr[:, 1, 2] = dg[:, 8]
    logr[:, 1, 0] = -dg[:, 6]; logr[:, 2, 0] = -dg[:, 7]; logr[:, 2, 1] = -dg[:, 8]
    S = np.empty((m, 3, 3))
    S[:, 0, 0] = dg[:, 0] + 1.0; S[:, 0, 1] = dg[:, 1]; S[:, 0, 2] = dg[:, 2]
    S[:, 1, 0] = dg[:, 1]; S[:, 1, 1] = dg[:, 3] + 1.0; S[:, 1, 2] = dg[:, 4]
    S[:, 2, 0] = dg[:, 2]; S[:, 2, 1] = dg[:, 4]; S[:, 2, 2] = dg[:, 5] + 1.0
    return rotation_exp(logr) @ S


def _rotation_log(R: np.ndarray) -> np.ndarray:
    """rotation_log_exp::log(Matrix3d) (utils_rotation.cpp:71-175) for one 3x3 matrix."""
    tol = _LOGEXP_TOL
    axis = np.zeros(3)
    angle = 0.0

    def _ret(angle, axis):
        t = np.zeros((3, 3))
        t[2, 1] = axis[0]; t[0, 2] = axis[1]; t[1, 0] = axis[2]
        return angle * (t - t.T)

    if np.linalg.norm(R.T @ R - np.eye(3)) > tol:
        # :73-77 returns with angle/axis uninitialised; not reachable for proper polar factors
        return _ret(0.0, axis)
    csin = (np.trace(R) - 1.0) / 2.0
    if csin < -1.0 or csin > 1.0:
        if abs(csin - 1.0) > tol and abs(csin + 1.0) > tol:
            return _ret(0.0, axis)
        csin = max(min(1.0, csin), -1.0)
    tangle = np.arccos(csin)
    if abs(tangle) < tol:
        return _ret(0.0, np.zeros(3))
    if abs(tangle - np.pi) < tol:
        B = (R + np.eye(3)) / 2.0
        k1 = np.sqrt(B[0, 0])
        k2 = np.sqrt(B[1, 1]) if k1 * B[0, 1] > 0.0 else -np.sqrt(B[1, 1])
        k3 = np.sqrt(B[2, 2]) if k1 * B[0, 2] > 0.0 else -np.sqrt(B[2, 2])
        return _ret(np.pi, np.array([k1, k2, k3]))
    taxis = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])

    def _check(tangle):
        t_axis = taxis / (2.0 * np.sin(tangle))
        sinv = np.sin(tangle)
        r01 = (1.0 - csin) * t_axis[0] * t_axis[1] - t_axis[2] * sinv
        r02 = (1.0 - csin) * t_axis[0] * t_axis[2] + t_axis[1] * sinv
        r10 = (1.0 - csin) * t_axis[0] * t_axis[1] + t_axis[2] * sinv
        r12 = (1.0 - csin) * t_axis[1] * t_axis[2] - t_axis[0] * sinv
        r20 = (1.0 - csin) * t_axis[0] * t_axis[2] - t_axis[1] * sinv
        r21 = (1.0 - csin) * t_axis[1] * t_axis[2] + t_axis[0] * sinv
        chk = ((R[0, 1] - r01) ** 2 + (R[0, 2] - r02) ** 2 + (R[1, 0] - r10) ** 2
               + (R[1, 2] - r12) ** 2 + (R[2, 0] - r20) ** 2 + (R[2, 1] - r21) ** 2)
        return chk, t_axis

    chk, t_axis = _check(tangle)
    if chk < tol:
        return _ret(tangle, t_axis)
    tangle = 2 * np.pi - tangle
    chk, t_axis = _check(tangle)
    return _ret(tangle, t_axis)


def _edge3(e1, e2, eps):
    """_getEdge3 lambda (deform_triangle_impl.hpp:152-161), batched.  Returns (ok, e3)."""
    e3 = np.cross(e1, e2)
    len1 = np.sqrt(np.einsum("ij,ij->i", e1, e1))
    len2 = np.sqrt(np.einsum("ij,ij->i", e2, e2))
    with np.errstate(divide="ignore", invalid="ignore"):
        abs_cos = np.abs(np.einsum("ij,ij->i", e1, e2) / (len1 * len2))
    ok = ~(abs_cos > (1.0 - eps))
    ok &= ~np.isnan(abs_cos) | True     # NaN > x is false in C++ too => "good"
    den = np.maximum(np.einsum("ij,ij->i", e3, e3) ** 0.25, eps)
    return ok, e3 / den[:, None]


def _triangle_mats(src, dst, tris, eps):
    a1, a2, a3 = (src[tris[:, i]].astype(np.float64) for i in range(3))
    b1, b2, b3 = (dst[tris[:, i]].astype(np.float64) for i in range(3))
    ea1, ea2, eb1, eb2 = a2 - a1, a3 - a1, b2 - b1, b3 - b1
    ok_a, ea3 = _edge3(ea1, ea2, eps)
    ok_b, eb3 = _edge3(eb1, eb2, eps)
    ok = ok_a & ok_b
    MA = np.stack([ea1, ea2, ea3], axis=2)        # columns = edges (:204-205)
    MB = np.stack([eb1, eb2, eb3], axis=2)
    T = np.tile(np.eye(3), (len(tris), 1, 1))
    if ok.any():
        T[ok] = MB[ok] @ np.linalg.inv(MA[ok])    # _getTransform :443-446
    return ok, T


class TriangleDeformationOracle:
    """Stateful restatement of deformation::TriangleDeformation + the pybind entry points."""

    def __init__(self):
        self.n_verts = 0
        self.n_tris = 0
        self.n_cnsts = 0
        self._lu = None

    # ----------------------------------------------------------------- set_target
    def set_target(self, verts, faces, cnsts=(), corrs=(), reg=1e-10) -> bool:
        V = _as_f32_verts(verts)
        F = _as_u32_tris(faces).astype(np.int64)
        c = np.ascontiguousarray(cnsts, dtype=np.uint32).astype(np.int64).reshape(-1)
        corr = np.ascontiguousarray(corrs, dtype=np.uint32).astype(np.int64).reshape(-1)
        n, m, k = len(V), len(F), len(c)
        blocks_per_tri = np.maximum(1, corr) if corr.size > 0 else np.ones(m, dtype=np.int64)  # :18-22
        assert len(blocks_per_tri) == m
        n_eq = int(blocks_per_tri.sum())
        # vertex <-> column maps (:36-73): free vertices keep their relative order,
        # constrained vertex c[i] -> column i of A_r
        is_c = np.zeros(n, dtype=bool)
        is_c[c] = True
        assert len(np.unique(c)) == k, "duplicate constraint index (reference asserts at :60)"
        col_to_vi_A = np.flatnonzero(~is_c)
        vi_to_col_A = np.full(n, -1, dtype=np.int64)
        vi_to_col_A[col_to_vi_A] = np.arange(n - k)
        vi_to_col_Ar = np.full(n, -1, dtype=np.int64)
        vi_to_col_Ar[c] = np.arange(k)
        U = triangle_frames(V, F)                                  # (m,2,3)
        # nine triplets per equation block (:102-117)
        tri_of_block = np.repeat(np.arange(m), blocks_per_tri)     # block k -> target triangle j
        Ub = U[tri_of_block]                                       # (n_eq,2,3)
        rows = (3 * np.arange(n_eq)[:, None] + np.arange(3)[None, :])          # (n_eq,3)
        coef = np.stack([-Ub[:, 0, :] - Ub[:, 1, :], Ub[:, 0, :], Ub[:, 1, :]], axis=1)  # (n_eq, corner, r)
        vidx = F[tri_of_block]                                     # (n_eq, corner)
        R = np.broadcast_to(rows[:, None, :], coef.shape).reshape(-1)
        VI = np.broadcast_to(vidx[:, :, None], coef.shape).reshape(-1)
        W = coef.reshape(-1)
        free = vi_to_col_A[VI] >= 0
        A = sp.csc_matrix((W[free], (R[free], vi_to_col_A[VI[free]])), shape=(3 * n_eq, n - k))
        Ar = sp.csc_matrix((W[~free], (R[~free], vi_to_col_Ar[VI[~free]])), shape=(3 * n_eq, max(k, 1)))
        At = A.T.tocsr()
        AtA = (At @ A).tocsc()
        if reg != 0:                                               # :126-131
            AtA = AtA + reg * sp.identity(n - k, format="csc")
        self.n_verts, self.n_tris, self.n_cnsts, self.n_eq = n, m, k, n_eq
        self.V, self.F = V, F
        self.A, self.Ar, self.At, self.AtA = A, Ar, At, AtA
        self.col_to_vi_A, self.col_to_vi_Ar = col_to_vi_A, c.copy()
        self.blocks_per_tri, self.tri_of_block = blocks_per_tri, tri_of_block
        self.U = U
        try:
            self._lu = spla.splu(AtA.tocsc())
        except RuntimeError:                                       # :134-139
            self._lu = None
            return False
        return True

    def is_same(self, num_verts, num_faces, num_cnsts) -> bool:    # pybind.cpp:119-126
        return (self.n_verts == num_verts and self.n_tris == num_faces and self.n_cnsts == num_cnsts)

    # ------------------------------------------------------------------- get_mesh
    def _solve_scatter(self, B, vert_cnsts):
        k = self.n_cnsts
        if k > 0:
            assert vert_cnsts is not None and len(vert_cnsts) > 0, "cnst_verts is not given"   # :274
            C32 = np.ascontiguousarray(vert_cnsts, dtype=np.float32).reshape(-1, 3)
            assert len(C32) == k
            B = B - self.Ar @ C32.astype(np.float64)               # :275-282
        X = self._lu.solve(np.asarray(self.At @ B))                # :286
        out = np.zeros((self.n_verts, 3), dtype=np.float32)
        out[self.col_to_vi_A] = X.astype(np.float32)               # :295-301
        if k > 0:
            out[self.col_to_vi_Ar] = C32                           # :302-308
        return out

    def get_mesh(self, deform_grad, vert_cnsts=(), corr_count=(), corr_faces=()):
        dg = np.ascontiguousarray(deform_grad, dtype=np.float64).reshape(-1, 9)
        cc = np.ascontiguousarray(corr_count, dtype=np.uint32).astype(np.int64).reshape(-1)
        cf = np.ascontiguousarray(corr_faces, dtype=np.uint32).astype(np.int64).reshape(-1)
        if cc.size == 0:                                           # :249-253 (pybind.cpp:113-114)
            T = dgrad_to_transforms(dg[: self.n_tris])
            B = np.transpose(T, (0, 2, 1)).reshape(-1, 3)          # rows 3i..3i+2 = T_i^T
        else:
            # :254-268 -- one block per (target tri, source tri); zero-corr triangles get identity.
            # fi advances by max(1, count) and corr_faces is indexed by that same block counter.
            blocks = np.maximum(1, cc)
            n_eq = int(blocks.sum())
            has = np.repeat(cc > 0, blocks)
            src = cf[:n_eq]
            T = np.tile(np.eye(3), (n_eq, 1, 1))
            if has.any():
                T[has] = dgrad_to_transforms(dg[src[has]])
            B = np.transpose(T, (0, 2, 1)).reshape(-1, 3)
        assert B.shape[0] == 3 * self.n_eq
        return self._solve_scatter(B, vert_cnsts)

    get_mesh_from_dg = get_mesh

    def get_mesh_from_dm(self, deform_mat, vert_cnsts=()):
        """deform_triangle_impl.hpp:391-397: each 9-vector is read column-major, i.e. the
        caller's row-major T lands transposed -- the same T^T block get_mesh builds."""
        dm = np.ascontiguousarray(deform_mat, dtype=np.float64).reshape(-1, 3, 3)
        B = np.transpose(dm[: self.n_tris], (0, 2, 1)).reshape(-1, 3)
        return self._solve_scatter(B, vert_cnsts)

    # ------------------------------------------------------------ get_deform_grad
    @staticmethod
    def get_deform_mat(verts_a, verts_b, faces, eps=1e-6):
        """deform_triangle_impl.hpp:313-380: row-major T per triangle, identity if degenerate."""
        Va, Vb = _as_f32_verts(verts_a), _as_f32_verts(verts_b)
        F = _as_u32_tris(faces).astype(np.int64)
        _, T = _triangle_mats(Va, Vb, F, eps)
        return T.reshape(-1).astype(np.float64)

    @staticmethod
    def get_deform_grad(verts_a, verts_b, faces, eps=1e-6):
        """deform_triangle_impl.hpp:144-213 + :448-470 (polar decomposition via SVD, log R)."""
        Va, Vb = _as_f32_verts(verts_a), _as_f32_verts(verts_b)
        F = _as_u32_tris(faces).astype(np.int64)
        ok, T = _triangle_mats(Va, Vb, F, eps)
        m = len(F)
        out = np.zeros((m, 9))
        if ok.any():
            Uu, s, Vt = np.linalg.svd(T[ok])
            Vv = np.transpose(Vt, (0, 2, 1))
            det = np.linalg.det(Uu @ Vt)
            temp = np.tile(np.eye(3), (len(s), 1, 1))
            temp[:, 2, 2] = det
            R = Uu @ temp @ Vt
            S = np.zeros((len(s), 3, 3))
            S[:, 0, 0], S[:, 1, 1], S[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
            scale = Vv @ temp @ S @ Vt
            logr = np.stack([_rotation_log(r) for r in R])
            g = np.stack([scale[:, 0, 0] - 1, scale[:, 0, 1], scale[:, 0, 2], scale[:, 1, 1] - 1,
                          scale[:, 1, 2], scale[:, 2, 2] - 1, logr[:, 0, 1], logr[:, 0, 2], logr[:, 1, 2]],
                         axis=1)
            out[ok] = g
        return out.reshape(-1)


# ----------------------------------------------------------------------------- glue
def pca_decode(coeff_scale, compT_scale, means_scale, coeff_rotat, compT_rotat, means_rotat,
               dtype=np.float32):
    """PcaInversion.forward x2 + data_to_anime_feat (output_module.py:115-116; model.py:246-257).

    coeff_* : (N,K*), compT_* : (out,K*), means_* : (out,).  Returns (N, n_tris*9) with per
    triangle [s00,s01,s02,s11,s12,s22,r01,r02,r12].  dtype float32 mirrors torch's fp32
    F.linear up to summation order; float64 gives the ground truth.
    """
    cs = np.asarray(coeff_scale, dtype=dtype); cr = np.asarray(coeff_rotat, dtype=dtype)
    scale = cs @ np.asarray(compT_scale, dtype=dtype).T + np.asarray(means_scale, dtype=dtype)
    rotat = cr @ np.asarray(compT_rotat, dtype=dtype).T + np.asarray(means_rotat, dtype=dtype)
    n = scale.shape[0]
    return np.concatenate([scale.reshape(n, -1, 6), rotat.reshape(n, -1, 3)], axis=2).reshape(n, -1)


def seek(ts, timestamps, sequence):
    """saber/data/stream/stream.py:20-46, restated: binary search + linear interpolation."""
    assert len(timestamps) == len(sequence)
    left, right = 0, len(timestamps)
    m = (left + right) // 2
    while left < right:
        m = (left + right) // 2
        tm = timestamps[m]
        tn = timestamps[m + 1] if m + 1 < len(timestamps) else ts + 1
        if tm <= ts < tn:
            break
        elif tm > ts:
            right = m
        else:
            left = m + 1
    if ts < timestamps[m] or ts > timestamps[-1]:
        return np.copy(sequence[m])
    if m + 1 >= len(timestamps):
        return np.copy(sequence[m])
    n = m + 1
    a = (timestamps[n] - ts) / (timestamps[n] - timestamps[m])
    return a * sequence[m] + (1 - a) * sequence[n]
