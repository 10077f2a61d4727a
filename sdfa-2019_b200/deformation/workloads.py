"""Synthetic inputs for the dgrad -> mesh path: templates, masks, dgrad sequences, PCA bases.

Pure numpy helpers shared by bench.py and the tests (no compute of the path itself).
Semantics follow the reference where it defines them:

* ``read_obj``          saber/data/mesh/io.py:23-68 (fan triangulation, 1-based -> 0-based)
* ``load_flame``        speech_anime/datasets/vocaset/template/FLAME_sample.obj +
                        speech_anime/datasets/vocaset/mask/non_face.py, shipped as the fixture
                        tests/golden/flame_template.npz (made by tests/golden/make_fixtures.py)
* ``subdivide``         SURVEY.md appendix B.4 (midpoint subdivision, mask = AND of the parents)
* dgrad layout          [N, n_tris, 9] = [s00,s01,s02,s11,s12,s22,r01,r02,r12] per triangle
                        (deform_triangle_impl.hpp:232-240; speech_anime/model/model.py:246-257)
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
FLAME_FIXTURE = os.path.join(REPO_ROOT, "tests", "golden", "flame_template.npz")

# scale/rotation PCA widths of the reference's dgrad model (speech_anime/config/model/dgrad.py:77-92)
K_SCALE, K_ROTAT = 85, 180


def read_obj(path, dtype=np.float32):
    verts, faces = [], []
    with open(path) as fp:
        for line in fp:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) for x in t[1:]]
                for i in range(len(idx) - 2):
                    faces.append((idx[0], idx[i + 1], idx[i + 2]))
    return np.asarray(verts, dtype=dtype), (np.asarray(faces, dtype=np.int64) - 1).astype(np.uint32)


def load_flame(path: str = FLAME_FIXTURE):
    """-> verts f32 (5023,3), faces u32 (9976,3), non_face_verts u32 (3762,), non_face_tris u32 (7375,)."""
    z = np.load(path)
    return z["verts"], z["faces"], z["non_face_verts"], z["non_face_tris"]


def bbox_diag(verts) -> float:
    v = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    return float(np.linalg.norm(v.max(0) - v.min(0)))


def subdivide(verts, faces, cnst_mask):
    """One midpoint subdivision.  New vertex per undirected edge, appended after the old ones;
    face (a,b,c) -> (a,ab,ca),(ab,b,bc),(ca,bc,c),(ab,bc,ca); midpoint constrained iff both ends are."""
    V = np.asarray(verts, dtype=np.float32).reshape(-1, 3)
    F = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    ue, inv = np.unique(e, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    m = len(F)
    ab, bc, ca = (len(V) + inv[:m], len(V) + inv[m:2 * m], len(V) + inv[2 * m:])
    mid = ((V[ue[:, 0]].astype(np.float64) + V[ue[:, 1]].astype(np.float64)) * 0.5).astype(np.float32)
    V2 = np.concatenate([V, mid], axis=0)
    a, b, c = F[:, 0], F[:, 1], F[:, 2]
    F2 = np.stack([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1),
                   np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)], axis=1).reshape(-1, 3)
    mask = np.asarray(cnst_mask, dtype=bool)
    mask2 = np.concatenate([mask, mask[ue[:, 0]] & mask[ue[:, 1]]])
    return V2, F2.astype(np.uint32), mask2


def flame_sub2():
    """Config 5 template: FLAME subdivided twice with the propagated default mask
    -> 79 936 v / 159 616 f / 59 283 constrained."""
    V, F, nfv, _ = load_flame()
    mask = np.zeros(len(V), dtype=bool)
    mask[nfv] = True
    for _ in range(2):
        V, F, mask = subdivide(V, F, mask)
    return V, F, np.flatnonzero(mask).astype(np.uint32)


def grid_mesh(nx=9, ny=7, seed=3, jitter=0.15):
    """Small open triangulated sheet with a little z relief: the tiny template for edge-case tests."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), indexing="ij")
    P = np.stack([xs, ys, 0.3 * np.sin(xs * 0.7) * np.cos(ys * 0.9)], axis=-1).reshape(-1, 3)
    P += jitter * rng.uniform(-1, 1, P.shape)
    P *= 0.01
    idx = np.arange(nx * ny).reshape(nx, ny)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=0)
    border = np.unique(np.concatenate([idx[0, :], idx[-1, :], idx[:, 0], idx[:, -1]]))
    return P.astype(np.float32), F.astype(np.uint32), border.astype(np.uint32)


def iid_dgrad(n_frames, n_tris, sigma=0.01, seed=0, start=0):
    """Config-1 style frames: N(0, sigma) iid in fp64 rounded to fp32.  Frame f of seed s is the
    same array whatever ``start``/``n_frames`` window is asked for (one generator stream per frame)."""
    out = np.empty((n_frames, n_tris * 9), dtype=np.float32)
    for i in range(n_frames):
        rng = np.random.default_rng([seed, start + i])
        out[i] = (sigma * rng.standard_normal(n_tris * 9)).astype(np.float32)
    return out


def smooth_displacement(verts, free_mask, amplitude, seed=0, n_waves=4):
    """Displace free vertices by a smooth (low-frequency sinusoidal) vector field of the given max
    amplitude; constrained vertices stay put.  Used to make *integrable* dgrads."""
    V = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    span = V.max(0) - V.min(0)
    d = np.zeros_like(V)
    for _ in range(n_waves):
        k = rng.uniform(0.3, 1.2, 3) * 2 * np.pi / span
        ph = rng.uniform(0, 2 * np.pi)
        direction = rng.standard_normal(3)
        d += np.sin(V @ k + ph)[:, None] * direction[None, :]
    d *= amplitude / np.abs(d).max()
    d[~np.asarray(free_mask, dtype=bool)] = 0.0
    return (V + d).astype(np.float32)


def random_pca(n_tris, seed=1, k_scale=K_SCALE, k_rotat=K_ROTAT, target_std=0.02, zero_tris=None):
    """Config-2 style random bases: orthonormal columns scaled so decoded dgrad std ~ target_std for
    N(0,1) coefficients; means = 0.001*N(0,1).  Rows of ``zero_tris`` are zeroed like the training
    data of the reference (datasets/vocaset/preload.py:778)."""
    rng = np.random.default_rng(seed)

    def basis(rows, k):
        q, _ = np.linalg.qr(rng.standard_normal((rows, k)))
        # a row of an orthonormal-column matrix has norm ~ sqrt(k/rows); scale rows to target_std
        return (q * (target_std * np.sqrt(rows / k))).astype(np.float32)

    cs, cr = basis(n_tris * 6, k_scale), basis(n_tris * 3, k_rotat)
    ms = (0.001 * rng.standard_normal(n_tris * 6)).astype(np.float32)
    mr = (0.001 * rng.standard_normal(n_tris * 3)).astype(np.float32)
    if zero_tris is not None and len(zero_tris):
        z = np.asarray(zero_tris, dtype=np.int64)
        cs.reshape(n_tris, 6, -1)[z] = 0; ms.reshape(n_tris, 6)[z] = 0
        cr.reshape(n_tris, 3, -1)[z] = 0; mr.reshape(n_tris, 3)[z] = 0
    return cs, ms, cr, mr


def random_coeffs(n_frames, seed=2, k_scale=K_SCALE, k_rotat=K_ROTAT):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_frames, k_scale)).astype(np.float32),
            rng.standard_normal((n_frames, k_rotat)).astype(np.float32))
