"""GPU tier: foreign-topology (correspondence) mode at FLAME size, fed from the files the reference's
``--mesh_constraints`` / ``--mesh_tricorres`` options name (viewer/frame.py:48-96, evaluate.sh:25-41), against the
compiled reference (deform_triangle_impl.hpp:18-22, 102-117, 246-269)."""
import numpy as np
import pytest

import deformation as D
from deformation import formats as FM
from deformation import workloads as W
from oracle import ref_loader
from oracle.dgrad_oracle import TriangleDeformationOracle

pytestmark = pytest.mark.gpu


def _files(tmp_path, V, F, nfv, n_src, seed):
    """Template OBJ + constraint list + triangle correspondences (0..3 source triangles per target triangle)."""
    rng = np.random.default_rng(seed)
    FM.write_obj(tmp_path / "target.obj", V, F)
    with open(tmp_path / "cnst.txt", "w") as fp:
        ids = [str(int(i)) for i in nfv]
        for a in range(0, len(ids), 40):
            fp.write(" ".join(ids[a:a + 40]) + "\n")
    counts = rng.choice([0, 1, 1, 1, 2, 3], size=len(F))
    recs = [(int(rng.integers(n_src)), d) for d in range(len(F)) for _ in range(counts[d])]
    order = rng.permutation(len(recs))                       # the file is not sorted by target triangle
    with open(tmp_path / "corr.txt", "w") as fp:
        fp.write(f"{len(recs)}\n")
        for i in order:
            fp.write(f"{recs[i][0]},{recs[i][1]},{rng.random():.4f}\n")
    return str(tmp_path / "target.obj"), str(tmp_path / "cnst.txt"), str(tmp_path / "corr.txt")


def test_flame_sized_correspondences_from_files(flame, tmp_path):
    import torch
    V0, F0, nfv0 = flame["V"], flame["F"], flame["nfv"]
    n_src = 7000                                             # the source mesh has another triangle count
    tpl, cpath, tpath = _files(tmp_path, V0, F0, nfv0, n_src, seed=11)
    V, F, c, corres = FM.load_template(tpl, cpath, tpath)
    assert np.array_equal(V, V0) and np.array_equal(F, F0) and np.array_equal(c, nfv0.astype(np.uint32))
    cc, cf = corres["corr_count"], corres["corr_faces"]
    tol = 1e-6 * W.bbox_diag(V)
    chk = ref_loader.RefSolver(1) if ref_loader.ref_available() else TriangleDeformationOracle()
    assert chk.set_target(V, F, cnsts=c, corrs=cc)
    rec = D.Reconstructor(V, F, cnsts=c, corrs=cc, device=0)
    assert rec.n_eq == int(np.maximum(cc, 1).sum())
    src = W.iid_dgrad(70, n_src, sigma=0.03, seed=12)
    Cm = (V[c] + np.float32(2e-4) * np.random.default_rng(13).standard_normal((len(c), 3)).astype(np.float32))
    # legacy single-frame call, the way frame.py:133-137 passes the correspondences on every call
    for i, C in ((0, V[c]), (1, Cm)):
        want = chk.get_mesh(src[i].astype(np.float64), vert_cnsts=C, corr_count=cc, corr_faces=cf)
        got = rec.get_mesh(src[i].astype(np.float64), vert_cnsts=C, corr_count=cc, corr_faces=cf)
        assert np.abs(got - want).max() <= tol, i
    # batched, device tensors, more than one 64-frame tile
    rec.set_constraint_positions(V[c])
    rec.set_correspondences(cc, cf, n_src_tris=n_src)
    out = rec.get_mesh_batch(torch.from_numpy(src).cuda()).cpu().numpy()
    for i in (0, 1, 63, 64, 69):
        want = chk.get_mesh(src[i].astype(np.float64), vert_cnsts=V[c], corr_count=cc, corr_faces=cf)
        assert np.abs(out[i] - want).max() <= tol, i
    # decode path: the PCA basis lives on the SOURCE topology (n_src triangles)
    cs, ms, cr, mr = W.random_pca(n_src, seed=14, k_scale=21, k_rotat=33)
    rec.set_pca(cs, ms, cr, mr)
    xs, xr = W.random_coeffs(5, seed=15, k_scale=21, k_rotat=33)
    out = rec.decode_and_get_mesh(xs, xr)
    s = torch.nn.functional.linear(torch.from_numpy(xs), torch.from_numpy(cs), torch.from_numpy(ms))
    r = torch.nn.functional.linear(torch.from_numpy(xr), torch.from_numpy(cr), torch.from_numpy(mr))
    dg = torch.cat((s.view(5, -1, 6), r.view(5, -1, 3)), dim=-1).view(5, -1).numpy()
    full = rec.decode_dgrad(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()).cpu().numpy()
    assert full.shape == (5, n_src * 9) and np.abs(full - dg).max() <= 2e-6
    for i in (0, 4):
        want = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[c], corr_count=cc, corr_faces=cf)
        assert np.abs(out[i] - want).max() <= tol, i
