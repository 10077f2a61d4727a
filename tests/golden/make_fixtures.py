"""Regenerates tests/golden/*.npz from the reference (run in the authoring container only).

    python tests/golden/make_fixtures.py            # needs /root/reference and `make -C oracle`

* flame_template.npz   the FLAME template + default mask the reference ships
                       (speech_anime/datasets/vocaset/template/FLAME_sample.obj read with
                       saber/data/mesh/io.py:23-68 semantics; datasets/vocaset/mask/non_face.py)
* golden_flame.npz     outputs of the UNMODIFIED reference module (oracle/_ref, built from
                       deformation/cpp/src by oracle/Makefile) on seeded inputs
* golden_small.npz     same for a tiny sheet mesh: unconstrained / correspondence / moved-constraint
                       modes, get_deform_grad, get_deform_mat, get_mesh_from_dm
The inputs are regenerated from seeds by deformation/workloads.py; only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "sdfa-2019_b200"))
sys.path.insert(0, ROOT)

REF = "/root/reference"


def make_template():
    from deformation import workloads as W
    V, F = W.read_obj(os.path.join(REF, "speech_anime/datasets/vocaset/template/FLAME_sample.obj"))
    ns = {}
    with open(os.path.join(REF, "speech_anime/datasets/vocaset/mask/non_face.py")) as fp:
        exec(fp.read(), ns)
    nfv = np.asarray(ns["non_face_verts"], dtype=np.uint32)
    nft = np.asarray(ns["non_face_tris"], dtype=np.uint32)
    assert V.shape == (5023, 3) and F.shape == (9976, 3) and len(nfv) == 3762 and len(nft) == 7375
    np.savez_compressed(os.path.join(HERE, "flame_template.npz"), verts=V, faces=F,
                        non_face_verts=nfv, non_face_tris=nft)


def make_flame_golden():
    from deformation import workloads as W
    from oracle.ref_loader import load_ref_module
    ref = load_ref_module()
    V, F, nfv, nft = W.load_flame()
    C = V[nfv]
    assert ref.set_target(verts=V, faces=F, cnsts=nfv)
    out = {}
    # config 1: iid sigma=0.01, seed 0; frames 0..119 -> keep 3 full frames + a float64 checksum of each
    dg = W.iid_dgrad(120, len(F), sigma=0.01, seed=0)
    verts = np.stack([ref.get_mesh(deform_grad=d.astype(np.float64), vert_cnsts=C) for d in dg])
    free = np.setdiff1d(np.arange(len(V)), nfv)
    out["iid_free_verts"] = verts[:, free][:8]                    # (8, 1261, 3)
    out["iid_checksum"] = verts.astype(np.float64).sum(axis=(1, 2))
    out["iid_abs_checksum"] = np.abs(verts.astype(np.float64) - V[None].astype(np.float64)).sum(axis=(1, 2))
    # KAT 1: zero dgrad -> template
    out["zero_verts_maxdiff"] = np.abs(ref.get_mesh(deform_grad=np.zeros(len(F) * 9), vert_cnsts=C) - V).max()
    # integrable deformations: 2 / 10 / 30 mm smooth displacement of the free vertices
    fm = np.ones(len(V), dtype=bool); fm[nfv] = False
    for amp_mm in (2, 10, 30):
        Vb = W.smooth_displacement(V, fm, amp_mm * 1e-3, seed=amp_mm)
        g = ref.get_deform_grad(verts_a=V, verts_b=Vb, faces=F)
        g32 = g.astype(np.float32)
        out[f"integ{amp_mm}_dgrad_checksum"] = np.float64(g.sum())
        out[f"integ{amp_mm}_dgrad_active"] = g32.reshape(-1, 9)[np.setdiff1d(np.arange(len(F)), nft)]
        back = ref.get_mesh(deform_grad=g32.astype(np.float64), vert_cnsts=C)
        out[f"integ{amp_mm}_free_verts"] = back[free]
        out[f"integ{amp_mm}_roundtrip_maxdiff"] = np.abs(back - Vb).max()
    # moved constraints (general A_r (c - c0) path): constrained verts translated + jittered
    rng = np.random.default_rng(11)
    C2 = (C + np.float32(0.002) + (1e-4 * rng.standard_normal(C.shape)).astype(np.float32)).astype(np.float32)
    out["moved_cnst_free_verts"] = ref.get_mesh(deform_grad=dg[0].astype(np.float64), vert_cnsts=C2)[free]
    out["is_same"] = np.array([ref.is_same(5023, 9976, 3762), ref.is_same(5023, 9976, 0)])
    np.savez_compressed(os.path.join(HERE, "golden_flame.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v), v if np.ndim(v) == 0 else "")


def degenerate_case(V, Vb, F):
    """Append three collinear vertices and one triangle over them."""
    extra = np.array([[0, 0, 0], [0.01, 0.01, 0.0], [0.02, 0.02, 0.0]], dtype=np.float32)
    Vd = np.concatenate([V, extra]); Vbd = np.concatenate([Vb, extra * np.float32(1.5)])
    Fd = np.concatenate([F, np.array([[len(V), len(V) + 1, len(V) + 2]], dtype=np.uint32)])
    return Vd, Vbd, Fd


def make_small_golden():
    from deformation import workloads as W
    from oracle.ref_loader import load_ref_module
    ref = load_ref_module()
    V, F, border = W.grid_mesh()
    m = len(F)
    out = {}
    dg = W.iid_dgrad(4, m, sigma=0.05, seed=7)
    # constrained border
    assert ref.set_target(verts=V, faces=F, cnsts=border)
    out["cnst_verts"] = np.stack([ref.get_mesh(deform_grad=d.astype(np.float64), vert_cnsts=V[border]) for d in dg])
    Cm = (V[border] + np.float32(0.001)).astype(np.float32)
    out["moved_cnst_verts"] = ref.get_mesh(deform_grad=dg[0].astype(np.float64), vert_cnsts=Cm)
    # single constraint
    assert ref.set_target(verts=V, faces=F, cnsts=np.array([5], dtype=np.uint32))
    out["one_cnst_verts"] = ref.get_mesh(deform_grad=dg[1].astype(np.float64), vert_cnsts=V[[5]])
    # unconstrained (gauge: translation), reg default
    assert ref.set_target(verts=V, faces=F)
    out["uncnst_verts"] = ref.get_mesh(deform_grad=dg[2].astype(np.float64))
    # correspondences: counts 0/1/2 per target triangle, sources into an 11-triangle "source" dgrad
    rng = np.random.default_rng(5)
    cc = rng.integers(0, 3, m).astype(np.uint32)
    cf = []
    for c in cc:
        cf += [0] if c == 0 else list(rng.integers(0, 11, c))
    cf = np.asarray(cf, dtype=np.uint32)
    src_dg = W.iid_dgrad(1, 11, sigma=0.05, seed=9)[0]
    assert ref.set_target(verts=V, faces=F, cnsts=border, corrs=cc)
    out["corr_count"], out["corr_faces"] = cc, cf
    out["corr_verts"] = ref.get_mesh(deform_grad=src_dg.astype(np.float64), vert_cnsts=V[border],
                                     corr_count=cc, corr_faces=cf)
    # inverse path
    fm = np.ones(len(V), dtype=bool); fm[border] = False
    Vb = W.smooth_displacement(V, fm, 0.004, seed=1)
    out["Vb"] = Vb
    out["deform_grad"] = ref.get_deform_grad(verts_a=V, verts_b=Vb, faces=F)
    out["deform_mat"] = ref.get_deform_mat(verts_a=V, verts_b=Vb, faces=F)
    assert ref.set_target(verts=V, faces=F, cnsts=border)
    out["from_dm_verts"] = ref.get_mesh_from_dm(deform_mat=out["deform_mat"], vert_cnsts=V[border])
    # degenerate (collinear) triangle in the inverse path -> zero grad / identity mat
    # (deform_triangle_impl.hpp:157-158).  NB a *repeated-vertex* triangle gives 0/0 = NaN there,
    # which passes the `> 1-eps` test and returns NaN gradients: undefined, not pinned.
    Vd, Vbd, Fd = degenerate_case(V, Vb, F)
    out["degen_deform_grad"] = ref.get_deform_grad(verts_a=Vd, verts_b=Vbd, faces=Fd)
    out["degen_deform_mat"] = ref.get_deform_mat(verts_a=Vd, verts_b=Vbd, faces=Fd)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    make_template()
    make_flame_golden()
    make_small_golden()
