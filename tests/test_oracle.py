"""Pins the CPU restatement (oracle/dgrad_oracle.py) against the reference: the committed golden
outputs of the unmodified reference module, and -- when oracle/_ref is built -- the live module.
The reference itself has no tests for this path, so these are the pins."""
import os

import numpy as np
import pytest

from deformation import workloads as W
from oracle.dgrad_oracle import TriangleDeformationOracle, pca_decode, seek
from oracle import ref_loader

ULP = 2.0 ** -23  # float32 relative spacing; positions are < 0.25 m so 1 ulp < 3e-8 m


@pytest.fixture(scope="module")
def flame_oracle(flame):
    o = TriangleDeformationOracle()
    assert o.set_target(flame["V"], flame["F"], cnsts=flame["nfv"])
    return o


def test_structure_matches_survey(flame_oracle, flame):
    o = flame_oracle
    assert (o.n_verts, o.n_tris, o.n_cnsts) == (5023, 9976, 3762)
    assert o.A.shape == (29928, 1261) and o.A.nnz == 22455
    assert o.Ar.nnz == 67329
    assert o.AtA.nnz == 8569
    active = np.unique(o.A.tocoo().row // 3)
    assert len(active) == 2601
    assert np.array_equal(np.setdiff1d(np.arange(9976), active), np.sort(flame["nft"]))
    assert abs(W.bbox_diag(flame["V"]) - 0.43998772) < 1e-6


def test_kat_zero_dgrad_is_template(flame_oracle, flame):
    V, nfv = flame["V"], flame["nfv"]
    out = flame_oracle.get_mesh(np.zeros(9976 * 9), vert_cnsts=V[nfv])
    assert np.abs(out - V).max() <= 4.7e-10 + 1e-12


def test_iid_frames_vs_golden(flame_oracle, flame, golden_flame):
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    dg = W.iid_dgrad(8, len(F), sigma=0.01, seed=0)
    for i in range(8):
        out = flame_oracle.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.array_equal(out[nfv], V[nfv])
        d = np.abs(out[free] - golden_flame["iid_free_verts"][i]).max()
        assert d <= 0.25 * ULP, d      # <= 1 ulp(float32) at |x| < 0.25
        assert abs(out.astype(np.float64).sum() - golden_flame["iid_checksum"][i]) < 1e-5


def test_integrable_vs_golden(flame_oracle, flame, golden_flame):
    V, F, nfv, nft = flame["V"], flame["F"], flame["nfv"], flame["nft"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    active = np.setdiff1d(np.arange(len(F)), nft)
    fm = np.ones(len(V), dtype=bool); fm[nfv] = False
    for amp in (2, 10, 30):
        Vb = W.smooth_displacement(V, fm, amp * 1e-3, seed=amp)
        g = flame_oracle.get_deform_grad(V, Vb, F)
        ga = golden_flame[f"integ{amp}_dgrad_active"]
        assert (np.abs(g.reshape(-1, 9)[active] - ga) <= 1e-6 * np.maximum(1.0, np.abs(ga))).all()
        assert abs(g.sum() - golden_flame[f"integ{amp}_dgrad_checksum"]) < 1e-6
        full = np.zeros((len(F), 9), dtype=np.float32); full[active] = ga
        out = flame_oracle.get_mesh(full.astype(np.float64).reshape(-1), vert_cnsts=V[nfv])
        assert np.abs(out[free] - golden_flame[f"integ{amp}_free_verts"]).max() <= 0.25 * ULP
        assert np.abs(out - Vb).max() < 2e-9      # KAT 2: get_deform_grad -> get_mesh round trip


def test_moved_constraints_vs_golden(flame_oracle, flame, golden_flame):
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    rng = np.random.default_rng(11)
    C = V[nfv]
    C2 = (C + np.float32(0.002) + (1e-4 * rng.standard_normal(C.shape)).astype(np.float32)).astype(np.float32)
    dg = W.iid_dgrad(1, len(F), sigma=0.01, seed=0)[0]
    out = flame_oracle.get_mesh(dg.astype(np.float64), vert_cnsts=C2)
    assert np.abs(out[free] - golden_flame["moved_cnst_free_verts"]).max() <= 0.25 * ULP
    assert np.array_equal(out[nfv], C2)


def test_is_same(flame_oracle, golden_flame):
    assert flame_oracle.is_same(5023, 9976, 3762) == bool(golden_flame["is_same"][0])
    assert flame_oracle.is_same(5023, 9976, 0) == bool(golden_flame["is_same"][1])


def test_small_mesh_modes_vs_golden(golden_small):
    V, F, border = W.grid_mesh()
    m = len(F)
    dg = W.iid_dgrad(4, m, sigma=0.05, seed=7)
    o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=border)
    for i in range(4):
        out = o.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[border])
        assert np.abs(out - golden_small["cnst_verts"][i]).max() < 1e-8
    Cm = (V[border] + np.float32(0.001)).astype(np.float32)
    assert np.abs(o.get_mesh(dg[0].astype(np.float64), vert_cnsts=Cm) - golden_small["moved_cnst_verts"]).max() < 1e-8
    assert np.abs(o.get_mesh_from_dm(golden_small["deform_mat"], vert_cnsts=V[border])
                  - golden_small["from_dm_verts"]).max() < 1e-8
    assert o.set_target(V, F, cnsts=np.array([5], dtype=np.uint32))
    assert np.abs(o.get_mesh(dg[1].astype(np.float64), vert_cnsts=V[[5]]) - golden_small["one_cnst_verts"]).max() < 2e-7
    # correspondences
    cc, cf = golden_small["corr_count"], golden_small["corr_faces"]
    src = W.iid_dgrad(1, 11, sigma=0.05, seed=9)[0]
    assert o.set_target(V, F, cnsts=border, corrs=cc)
    out = o.get_mesh(src.astype(np.float64), vert_cnsts=V[border], corr_count=cc, corr_faces=cf)
    assert np.abs(out - golden_small["corr_verts"]).max() < 1e-8


def test_small_mesh_unconstrained_modulo_translation(golden_small):
    """No constraints: the system is singular up to a translation; compare centred."""
    V, F, _ = W.grid_mesh()
    dg = W.iid_dgrad(4, len(F), sigma=0.05, seed=7)
    o = TriangleDeformationOracle()
    assert o.set_target(V, F)
    out = o.get_mesh(dg[2].astype(np.float64))
    ref = golden_small["uncnst_verts"]
    assert np.abs((out - out.mean(0)) - (ref - ref.mean(0))).max() < 1e-6


def test_inverse_path_vs_golden(golden_small):
    V, F, _ = W.grid_mesh()
    Vb = golden_small["Vb"]
    o = TriangleDeformationOracle()
    assert np.abs(o.get_deform_grad(V, Vb, F) - golden_small["deform_grad"]).max() < 1e-9
    assert np.abs(o.get_deform_mat(V, Vb, F) - golden_small["deform_mat"]).max() < 1e-9
    from tests.golden.make_fixtures import degenerate_case
    Vd, Vbd, Fd = degenerate_case(V, Vb, F)
    g = o.get_deform_grad(Vd, Vbd, Fd)
    assert np.abs(g - golden_small["degen_deform_grad"]).max() < 1e-9
    assert np.all(g.reshape(-1, 9)[-1] == 0)
    t = o.get_deform_mat(Vd, Vbd, Fd)
    assert np.abs(t - golden_small["degen_deform_mat"]).max() < 1e-9
    assert np.array_equal(t.reshape(-1, 3, 3)[-1], np.eye(3))


def test_pca_decode_layout():
    rng = np.random.default_rng(0)
    nt = 5
    cs, ms, cr, mr = W.random_pca(nt, seed=1, k_scale=4, k_rotat=3)
    xs, xr = rng.standard_normal((2, 4)).astype(np.float32), rng.standard_normal((2, 3)).astype(np.float32)
    d = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64).reshape(2, nt, 9)
    s = (xs.astype(np.float64) @ cs.T.astype(np.float64) + ms).reshape(2, nt, 6)
    r = (xr.astype(np.float64) @ cr.T.astype(np.float64) + mr).reshape(2, nt, 3)
    assert np.allclose(d[:, :, :6], s) and np.allclose(d[:, :, 6:], r)


def test_seek_interpolation():
    ts = [0.0, 1.0, 2.0]
    seq = np.array([[0.0], [10.0], [30.0]])
    assert seek(0.5, ts, seq)[0] == 5.0
    assert seek(1.25, ts, seq)[0] == 15.0
    assert seek(5.0, ts, seq)[0] == 30.0
    assert seek(-1.0, ts, seq)[0] == 0.0


@pytest.mark.skipif(not os.path.exists("/root/reference/saber/data/stream/stream.py"),
                    reason="reference tree not present (GPU box)")
def test_seek_vs_reference_source():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_stream", "/root/reference/saber/data/stream/stream.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    ts = np.cumsum(rng.uniform(0.01, 0.03, 50))
    seq = rng.standard_normal((50, 7))
    for q in rng.uniform(ts[0] - 0.05, ts[-1] + 0.05, 200):
        assert np.array_equal(seek(q, ts, seq), mod.seek(q, ts, seq))


@pytest.mark.skipif(not ref_loader.ref_available(), reason="oracle/_ref not built")
def test_restatement_vs_live_reference(flame):
    """The restatement against the unmodified reference compiled from its own sources."""
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    ref = ref_loader.load_ref_module()
    assert ref.set_target(verts=V, faces=F, cnsts=nfv)
    o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=nfv)
    dg = W.iid_dgrad(3, len(F), sigma=0.05, seed=21)
    for d in dg:
        a = ref.get_mesh(deform_grad=d.astype(np.float64), vert_cnsts=V[nfv])
        b = o.get_mesh(d.astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(a - b).max() <= 0.25 * ULP
    # shim (threaded CPU baseline) == pybind module
    rs = ref_loader.RefSolver(2)
    assert rs.set_target(V, F, cnsts=nfv)
    outs, secs = rs.get_mesh_batch(dg, V[nfv])
    assert secs > 0
    for i, d in enumerate(dg):
        assert np.array_equal(outs[i], ref.get_mesh(deform_grad=d.astype(np.float64), vert_cnsts=V[nfv]))
