"""GPU tier (pytest -m gpu, run on a B200 through gpurun): the CUDA path, called through the C ABI /
the drop-in ``deformation`` package, against the oracle on identical inputs.

Tolerance (BASELINE.json north_star): max per-vertex deviation <= 1e-6 x template bounding-box diagonal
(4.4e-7 m for FLAME).  The checker is the compiled reference (oracle/_ref) when it travelled with the
snapshot, else the numpy restatement (oracle/dgrad_oracle.py); committed golden outputs are checked too."""
import numpy as np
import pytest

import deformation as D
from deformation import workloads as W
from oracle import ref_loader
from oracle.dgrad_oracle import TriangleDeformationOracle, pca_decode

pytestmark = pytest.mark.gpu


def _checker(V, F, cnsts=(), corrs=()):
    if ref_loader.ref_available():
        o = ref_loader.RefSolver(1)
    else:
        o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=cnsts, corrs=corrs)
    return o


@pytest.fixture(scope="module")
def rec(flame):
    return D.Reconstructor(flame["V"], flame["F"], cnsts=flame["nfv"], device=0)


@pytest.fixture(scope="module")
def chk(flame):
    return _checker(flame["V"], flame["F"], flame["nfv"])


def test_native_library_is_what_runs(rec):
    before = D.lib.sdfa_launch_count()
    rec.get_mesh_batch(np.zeros((1, 9976 * 9), dtype=np.float32))
    assert D.lib.sdfa_launch_count() >= before + 3      # assembly + solve + fill kernels


def test_kat_zero_dgrad_is_template(rec, flame):
    V, nfv = flame["V"], flame["nfv"]
    out = rec.get_mesh(np.zeros(9976 * 9), vert_cnsts=V[nfv])
    assert out.dtype == np.float32 and out.shape == (5023, 3)
    assert np.abs(out - V).max() <= 1e-9
    assert np.array_equal(out[nfv], V[nfv])


def test_config1_120_frames_vs_golden_and_checker(rec, chk, flame, golden_flame):
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    dg = W.iid_dgrad(120, len(F), sigma=0.01, seed=0)
    out = rec.get_mesh_batch(dg)
    assert out.shape == (120, 5023, 3)
    assert np.abs(out[:8, free] - golden_flame["iid_free_verts"]).max() <= tol
    assert np.abs(out.astype(np.float64).sum(axis=(1, 2)) - golden_flame["iid_checksum"]).max() < 5023 * 3 * tol
    for i in (0, 31, 32, 64, 119):
        ref = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(out[i] - ref).max() <= tol
        assert np.array_equal(out[i][nfv], V[nfv])
    # the legacy one-frame float64 entry point gives the same vertices as the batch
    one = rec.get_mesh(dg[5].astype(np.float64), vert_cnsts=V[nfv])
    assert np.array_equal(one, out[5])


def test_integrable_deformations_vs_golden(rec, flame, golden_flame):
    V, F, nfv, nft, tol = flame["V"], flame["F"], flame["nfv"], flame["nft"], flame["tol"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    active = np.setdiff1d(np.arange(len(F)), nft)
    for amp in (2, 10, 30):
        full = np.zeros((len(F), 9), dtype=np.float32)
        full[active] = golden_flame[f"integ{amp}_dgrad_active"]
        out = rec.get_mesh(full.reshape(-1).astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(out[free] - golden_flame[f"integ{amp}_free_verts"]).max() <= tol, amp


def test_large_sigma_and_partial_tiles(rec, chk, flame):
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    for n in (1, 31, 33, 97):
        dg = W.iid_dgrad(n, len(F), sigma=0.2, seed=100 + n)
        out = rec.get_mesh_batch(dg)
        for i in sorted({0, n // 2, n - 1}):
            ref = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
            assert np.abs(out[i] - ref).max() <= tol, (n, i)
    assert rec.get_mesh_batch(np.zeros((0, len(F) * 9), dtype=np.float32)).shape == (0, 5023, 3)


def test_moved_constraints(rec, flame, golden_flame):
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    rng = np.random.default_rng(11)
    C = V[nfv]
    C2 = (C + np.float32(0.002) + (1e-4 * rng.standard_normal(C.shape)).astype(np.float32)).astype(np.float32)
    dg = W.iid_dgrad(1, len(F), sigma=0.01, seed=0)[0]
    out = rec.get_mesh(dg.astype(np.float64), vert_cnsts=C2)
    assert np.abs(out[free] - golden_flame["moved_cnst_free_verts"]).max() <= tol
    assert np.array_equal(out[nfv], C2)
    out = rec.get_mesh(dg.astype(np.float64), vert_cnsts=C)      # and back to the template positions
    assert np.abs(out[free] - golden_flame["iid_free_verts"][0]).max() <= tol


def test_torch_device_tensors_and_linearity_of_batching(rec, flame):
    import torch
    F = flame["F"]
    dg = torch.from_numpy(W.iid_dgrad(200, len(F), sigma=0.03, seed=5)).cuda()
    out = rec.get_mesh_batch(dg)
    assert out.is_cuda and out.shape == (200, 5023, 3)
    # size-independent property: a frame's result does not depend on its batch or tile position
    perm = torch.randperm(200, generator=torch.Generator().manual_seed(0)).cuda()
    out2 = rec.get_mesh_batch(dg[perm].contiguous())
    assert torch.equal(out2, out[perm])
    # strided input rows
    wide = torch.zeros((50, len(F) * 9 + 64), dtype=torch.float32, device="cuda")
    wide[:, : len(F) * 9] = dg[:50]
    out3 = rec.get_mesh_batch(wide[:, : len(F) * 9])
    assert torch.equal(out3, out[:50])


def test_dropin_singleton_call_pattern(flame, golden_flame):
    """The exact sequence of speech_anime/viewer/frame.py:42,118-137."""
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    assert D.set_target(verts=np.reshape(V, (-1, 3)), faces=np.reshape(F, (-1, 3)), cnsts=list(nfv), corrs=[]) is True
    assert D.is_same(V.shape[0], F.shape[0], len(nfv))
    assert not D.is_same(V.shape[0], F.shape[0], 0)
    frame = W.iid_dgrad(1, len(F), sigma=0.01, seed=0)[0]
    out = D.get_mesh(deform_grad=frame.astype(np.float64), vert_cnsts=np.reshape(V, (-1, 3))[nfv],
                     corr_count=[], corr_faces=[])
    assert out.shape == (5023, 3) and out.dtype == np.float32
    assert np.abs(out[free] - golden_flame["iid_free_verts"][0]).max() <= tol
    assert np.array_equal(D.get_mesh_from_dg(deform_grad=frame.astype(np.float64), vert_cnsts=V[nfv]), out)


def test_small_mesh_modes_vs_golden(golden_small):
    V, F, border = W.grid_mesh()
    tol = 1e-6 * W.bbox_diag(V)
    dg = W.iid_dgrad(4, len(F), sigma=0.05, seed=7)
    r = D.Reconstructor(V, F, cnsts=border, device=0)
    assert np.abs(r.get_mesh_batch(dg) - golden_small["cnst_verts"]).max() <= tol
    assert np.abs(r.get_mesh_from_dm(golden_small["deform_mat"], vert_cnsts=V[border])
                  - golden_small["from_dm_verts"]).max() <= tol
    Cm = (V[border] + np.float32(0.001)).astype(np.float32)
    assert np.abs(r.get_mesh(dg[0].astype(np.float64), vert_cnsts=Cm) - golden_small["moved_cnst_verts"]).max() <= tol
    with pytest.raises(D.SdfaError):
        r.get_mesh(dg[0].astype(np.float64))                      # constraints but no vert_cnsts (impl.hpp:274)
    # correspondences (frame.py passes them on every call)
    cc, cf = golden_small["corr_count"], golden_small["corr_faces"]
    rc = D.Reconstructor(V, F, cnsts=border, corrs=cc, device=0)
    src = W.iid_dgrad(1, 11, sigma=0.05, seed=9)[0]
    out = rc.get_mesh(src.astype(np.float64), vert_cnsts=V[border], corr_count=cc, corr_faces=cf)
    assert np.abs(out - golden_small["corr_verts"]).max() <= tol
    # unconstrained: defined modulo a translation; the reference's own translation is
    # fp64 rounding noise amplified by 1/reg, ours pins one vertex, so compare centred
    ru = D.Reconstructor(V, F, device=0)
    out = ru.get_mesh(dg[2].astype(np.float64))
    ref = golden_small["uncnst_verts"]
    assert np.abs((out - out.mean(0)) - (ref - ref.mean(0))).max() <= 2e-6
    assert np.abs(out).max() < 1.0            # and it stays near the template instead of drifting off


def test_tensor_core_decode_matches_fp64(rec, flame):
    """K1 (tcgen05, 3xTF32) vs the float64 ground truth, against the error torch's own fp32 F.linear makes."""
    import torch
    F, nft = flame["F"], flame["nft"]
    cs, ms, cr, mr = W.random_pca(len(F), seed=1)
    rec.set_pca(cs, ms, cr, mr)
    for n in (1, 127, 300):
        xs, xr = W.random_coeffs(n, seed=40 + n)
        got = rec.decode_compact(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()).cpu().numpy()
        lay = rec.compact_layout()
        used = lay >= 0
        assert got.shape == (n, len(lay)) and len(np.unique(lay[used] // 9)) == 2601
        got = got[:, used]
        dg64 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64)[:, lay[used]]
        dg32 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float32)[:, lay[used]]
        err, err32 = np.abs(got - dg64).max(), np.abs(dg32 - dg64).max()
        # 3xTF32 is a few times coarser than fp32 FMA accumulation (operands truncated to 2 x 11 bits,
        # tensor-core accumulation) but must stay well inside what 4.4e-7 m vertex parity needs:
        # |dgrad error| <~ 1e-6
        assert err <= 5e-7, (n, err, err32)


def test_tensor_core_decode_other_basis_sizes(flame):
    """K1 with bases of other widths than the reference's 85 + 180, on both generations of the kernel: the FP16-split
    kernel (both frames operands resident: up to five 64-wide K-blocks over the two parts, partial last blocks, widths
    that are exact multiples of a block so that the means' column opens a new one) and the TF32-split kernel it falls
    back to for wider bases (forced with ``decode=tf32`` for the narrow ones: streamed scale operand, one-block bases).
    Several frame tiles, so that CTA pairs change tiles mid-walk."""
    import torch
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    for opts, want_kind, sizes in (({}, 16, ((120, 20), (96, 31), (32, 64), (5, 200), (63, 127))),
                                   ({}, 32, ((300, 40),)),                       # too wide for the resident operands
                                   ({"decode": "tf32"}, 32, ((120, 20), (96, 31), (32, 64), (5, 200)))):
        r = D.Reconstructor(V, F, cnsts=nfv, device=0, options=opts)
        for ks, kr in sizes:
            cs, ms, cr, mr = W.random_pca(len(F), seed=3, k_scale=ks, k_rotat=kr)
            r.set_pca(cs, ms, cr, mr)
            assert int(r.debug("decode_kind")[0]) == want_kind, (opts, ks, kr)
            n = 700
            xs, xr = W.random_coeffs(n, seed=ks, k_scale=ks, k_rotat=kr)
            xs[3] *= 1e-4                                     # per-frame scaling: a frame of tiny and one of huge coefficients
            xr[5] *= 300.0
            got = r.decode_compact(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()).cpu().numpy()
            lay = r.compact_layout()
            used = lay >= 0
            dg64 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64)[:, lay[used]]
            err = np.abs(got[:, used] - dg64)
            big = np.zeros(n, dtype=bool)
            big[5] = True
            assert err[~big].max() <= 5e-7, (opts, ks, kr)
            assert err[big].max() <= 5e-7 * 300, (opts, ks, kr)   # relative to that frame's magnitude
        r.close()


def test_decode_and_reconstruct_config2(rec, chk, flame):
    """Config 2: 240 frames of PCA coefficients -> vertices; oracle = fp32 F.linear + cat + reference get_mesh."""
    import torch
    V, F, nfv, nft, tol = flame["V"], flame["F"], flame["nfv"], flame["nft"], flame["tol"]
    cs, ms, cr, mr = W.random_pca(len(F), seed=1)
    xs, xr = W.random_coeffs(240, seed=2)
    rec.set_pca(cs, ms, cr, mr)
    out = rec.decode_and_get_mesh(xs, xr)
    assert out.shape == (240, 5023, 3)
    # decode alone: the tensor data_to_anime_feat returns (model.py:246-257), vs torch CPU fp32 F.linear
    dg_dev = rec.decode_dgrad(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()).cpu().numpy()
    lin_s = torch.nn.functional.linear(torch.from_numpy(xs), torch.from_numpy(cs), torch.from_numpy(ms))
    lin_r = torch.nn.functional.linear(torch.from_numpy(xr), torch.from_numpy(cr), torch.from_numpy(mr))
    dg_ref = torch.cat((lin_s.view(240, -1, 6), lin_r.view(240, -1, 3)), dim=-1).view(240, -1).numpy()
    dg64 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64)
    assert np.abs(dg_dev - dg64).max() <= 2 * np.abs(dg_ref - dg64).max() + 1e-7
    for i in (0, 100, 239):
        ref = chk.get_mesh(dg_ref[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(out[i] - ref).max() <= tol
    # torch path == numpy path
    out_t = rec.decode_and_get_mesh(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda())
    assert np.array_equal(out_t.cpu().numpy(), out)


def test_inverse_path_vs_golden_and_round_trip(flame, golden_flame, golden_small):
    """mesh -> dgrad on the GPU (getDeformationGradients, impl.hpp:144-213) and the mesh -> dgrad -> mesh
    round trip of SURVEY section 4 (KAT 2)."""
    import torch
    from tests.golden.make_fixtures import degenerate_case
    V, F, border = W.grid_mesh()
    Vb = golden_small["Vb"]
    assert np.abs(D.get_deform_grad(V, Vb, F) - golden_small["deform_grad"]).max() < 1e-9
    assert np.abs(D.get_deform_mat(verts_a=V, verts_b=Vb, faces=F, eps=1e-6) - golden_small["deform_mat"]).max() < 1e-9
    Vd, Vbd, Fd = degenerate_case(V, Vb, F)
    g = D.get_deform_grad(Vd, Vbd, Fd)
    assert np.abs(g - golden_small["degen_deform_grad"]).max() < 1e-9 and np.all(g.reshape(-1, 9)[-1] == 0)
    assert np.array_equal(D.get_deform_mat(Vd, Vbd, Fd).reshape(-1, 3, 3)[-1], np.eye(3))
    # FLAME: integrable sets against the reference's own get_deform_grad output, then back to the mesh
    Vf, Ff, nfv, nft, tol = flame["V"], flame["F"], flame["nfv"], flame["nft"], flame["tol"]
    active = np.setdiff1d(np.arange(len(Ff)), nft)
    fm = np.ones(len(Vf), dtype=bool); fm[nfv] = False
    targets = np.stack([W.smooth_displacement(Vf, fm, a * 1e-3, seed=a) for a in (2, 10, 30)])
    dg = D.get_deform_grad_batch(Vf, targets, Ff)
    assert dg.shape == (3, len(Ff) * 9) and dg.dtype == torch.float32
    for k, amp in enumerate((2, 10, 30)):
        ga = golden_flame[f"integ{amp}_dgrad_active"]
        got = dg[k].cpu().numpy().reshape(-1, 9)[active]
        assert (np.abs(got - ga) <= 2e-6 * np.maximum(1.0, np.abs(ga))).all(), amp
    rec = D.Reconstructor(Vf, Ff, cnsts=nfv, device=0)
    back = rec.get_mesh_batch(dg).cpu().numpy()
    assert np.abs(back - targets).max() <= tol


def test_config5_subdivided_template():
    """Config 5: FLAME subdivided twice (79 936 v / 159 616 f, 20 653 unknowns, nnz(L) ~ 7.8e5): the factor's
    resident rows only fit in shared memory with 16 frames per tile.  Parity on the same 1e-6 x bbox bar."""
    V, F, c = W.flame_sub2()
    assert V.shape == (79936, 3) and F.shape == (159616, 3) and len(c) == 59283
    tol = 1e-6 * W.bbox_diag(V)
    r = D.Reconstructor(V, F, cnsts=c, device=0)
    assert r.n_free == 20653 and int(r.debug("stats")[14]) in (8, 16)
    o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=c)
    dg = W.iid_dgrad(37, len(F), sigma=0.01, seed=5)
    out = r.get_mesh_batch(dg)
    for i in (0, 15, 16, 36):
        ref = o.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[c])
        assert np.abs(out[i] - ref).max() <= tol, i
        assert np.array_equal(out[i][c], V[c])
    # KAT: zero dgrad gives back the template
    z = r.get_mesh_batch(np.zeros((1, len(F) * 9), dtype=np.float32))
    assert np.abs(z[0] - V).max() <= 1e-8


def test_decode_path_on_other_templates():
    """The staged path (K1 -> k_assemble -> solve -> output) on templates other than FLAME -- a 63-vertex grid (row
    blocks with fewer equations than warps, tensor plan of a tiny system) and the subdivided FLAME of config 5 (179 row
    blocks, SIMT solve) -- against the gather path fed with the same decoded dgrad, which the tests above pin to the
    reference."""
    import torch
    Vg, Fg, border = W.grid_mesh()
    Vs, Fs, cs_ = W.flame_sub2()
    for V, F, c in ((Vg, Fg, border), (Vs, Fs, cs_)):
        tol = 1e-6 * W.bbox_diag(V)
        r = D.Reconstructor(V, F, cnsts=c, device=0)
        cs, ms, cr, mr = W.random_pca(len(F), seed=4, k_scale=20, k_rotat=30)
        r.set_pca(cs, ms, cr, mr)
        xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(70, seed=6, k_scale=20, k_rotat=30))
        out = r.decode_and_get_mesh(xs, xr)
        ref = r.get_mesh_batch(r.decode_dgrad(xs, xr))
        assert out.shape == (70, len(V), 3)
        assert float((out - ref).abs().max()) <= 0.25 * tol, len(V)
        assert torch.equal(out[:, torch.as_tensor(np.asarray(c), device="cuda").long()],
                           torch.from_numpy(V[np.asarray(c)]).cuda().expand(70, -1, -1))
        r.close()


def test_tensor_and_simt_solvers_agree(chk, flame):
    """K3T (tcgen05 block products, operands in tensor memory) and K3 (SIMT sweeps) on the same frames, against each
    other and the checker; frame counts around the 128-column tile boundaries of K3T."""
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    rt = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="tensor")
    rs = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="simt")
    assert rt.debug("ts_stats")[0] == 1 and rs.debug("ts_stats")[0] == 0
    dg = W.iid_dgrad(300, len(F), sigma=0.1, seed=11)
    for n in (1, 127, 129, 300):
        a, b = rt.get_mesh_batch(dg[:n]), rs.get_mesh_batch(dg[:n])
        assert not np.isnan(a).any()
        assert np.abs(a - b).max() <= 0.2 * tol, n
    a = rt.get_mesh_batch(dg)
    for i in (0, 127, 128, 255, 256, 299):
        ref = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(a[i] - ref).max() <= tol, i
    # a frame's result does not depend on what else is in the batch or where it sits in a tile
    assert np.array_equal(rt.get_mesh_batch(dg[200:201])[0], a[200])


def test_tensor_plan_variants_are_bit_equal(flame):
    """The planner's alternatives of the K3T plan -- one epilogue stream instead of two, no separate read events -- move
    the same values through the same products in the same order: bit-equal results, over many tiles per CTA."""
    import os
    import torch
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.05, seed=31)).cuda().repeat(400, 1)      # 25 600 frames: 600 tiles
    outs = []
    for env in ({}, {"SDFA_TS_STREAMS": "1"}, {"SDFA_TS_EARLY": "0"}):
        os.environ.update(env)
        try:
            r = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="tensor")
        finally:
            for k in env:
                del os.environ[k]
        assert int(r.debug("ts_stats")[14]) == (1 if env.get("SDFA_TS_STREAMS") == "1" else 2)
        outs.append(r.get_mesh_batch(dg))
        r.close()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(outs[0][:64], outs[0][-64:])            # every tile of every persistent CTA


def test_config3_batch_sizes_give_identical_frames(rec, flame):
    """Config 3 (solve sweep over RHS batches 64..4096): the same frames reconstructed in batches of different size
    are bit-identical -- tiles are independent and the summation orders are fixed (no atomics anywhere)."""
    import torch
    F = flame["F"]
    base = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.02, seed=21)).cuda()
    big = base.repeat(64, 1)                                   # 4096 frames
    ref = rec.get_mesh_batch(base)
    for n in (64, 128, 256, 512, 1024, 2048, 4096):
        out = rec.get_mesh_batch(big[:n])
        assert torch.equal(out[:64], ref) and torch.equal(out[n - 64:], ref), n


def test_chunk_pipeline_is_bit_identical(rec, flame):
    """Large batches run chunk by chunk with the output kernel of a chunk on a second stream (api.cpp
    reconstruct_chunks): forced to 256-frame chunks, both entry points return exactly what the single pass returns,
    ragged last chunk included, and back-to-back calls do not race on the two scratch buffers."""
    import torch
    F = flame["F"]
    n = 256 * 5 + 77
    dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.02, seed=33)).cuda().repeat((n + 63) // 64, 1)[:n].contiguous()
    cs, ms, cr, mr = W.random_pca(len(F), seed=1)
    rec.set_pca(cs, ms, cr, mr)
    xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(n, seed=8))
    rec.set_option("pipe_chunk", 0)
    ref_a, ref_b = rec.get_mesh_batch(dg), rec.decode_and_get_mesh(xs, xr)
    rec.set_option("pipe_chunk", 256)
    for _ in range(3):
        out_a, out_b = rec.get_mesh_batch(dg), rec.decode_and_get_mesh(xs, xr)
        assert torch.equal(out_a, ref_a) and torch.equal(out_b, ref_b)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        out_c = rec.decode_and_get_mesh(xs, xr, stream=side.cuda_stream)
    side.synchronize()
    assert torch.equal(out_c, ref_b)
    rec.set_option("pipe_chunk", -1)                  # back to the automatic chunk size (the handle is shared)


def test_config4_network_to_mesh_on_device(rec, chk, flame):
    """Config 4 in miniature: synthetic audio -> mel + deltas -> the reference's network architecture, random init
    (deformation/frontend.py restates speech_anime/config/model/dgrad.py:56-100 in plain torch) -> PCA coefficients ->
    K1..K5, everything staying on the device; checked against fp32 F.linear decode + the reference solver on the
    network's own coefficients."""
    import torch
    from deformation import frontend as FE
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    cs, ms, cr, mr = W.random_pca(len(F), seed=1)
    rec.set_pca(cs, ms, cr, mr)
    dev = torch.device("cuda", 0)
    audio = FE.band_limited_noise(2, seconds=2.0, seed=4, device=dev)
    with torch.no_grad():
        feats = FE.MelFeatures().to(dev)(audio, 120)                 # [2, 120, 64, 128, 3]
        net = FE.build_network(seed=0, device=dev, coeff_gain=8.0)
        spk = torch.tensor([1, 6], device=dev).repeat_interleave(120)
        xs, xr = net(feats.reshape(-1, 64, 128, 3), spk)
    assert xs.shape == (240, 85) and xr.shape == (240, 180) and float(xs.std()) > 0.05
    out = rec.decode_and_get_mesh(xs.contiguous(), xr.contiguous())
    assert out.is_cuda and out.shape == (240, 5023, 3)
    lin_s = torch.nn.functional.linear(xs.cpu(), torch.from_numpy(cs), torch.from_numpy(ms))
    lin_r = torch.nn.functional.linear(xr.cpu(), torch.from_numpy(cr), torch.from_numpy(mr))
    dg = torch.cat((lin_s.view(240, -1, 6), lin_r.view(240, -1, 3)), dim=-1).view(240, -1).numpy()
    for i in (0, 119, 120, 239):
        ref = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(out[i].cpu().numpy() - ref).max() <= tol, i


def test_seek_batch_matches_stream_seek(rec, chk, flame):
    """§8(f) rank 2: saber.stream.seek batched on the device -- bit-equal to the restated reference per timestamp
    (in range, exact hits, before the first and after the last sample), and seeking coefficients then decoding
    equals decoding then seeking the dgrad."""
    import torch
    from oracle.dgrad_oracle import seek
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    rng = np.random.default_rng(5)
    ts = np.cumsum(rng.uniform(0.01, 0.03, 50))                 # uneven sampling times
    seq = rng.normal(size=(50, 7, 3)).astype(np.float32)
    q = np.concatenate([rng.uniform(ts[0] - 0.05, ts[-1] + 0.05, 200), ts[[0, 17, 49]], [ts[0] - 1.0, ts[-1] + 1.0]])
    got = D.seek_batch(q, ts, torch.from_numpy(seq).cuda()).cpu().numpy()
    for i, t in enumerate(q):
        assert np.array_equal(got[i], np.asarray(seek(t, ts, seq)).astype(np.float32)), (i, t)
    # coefficients at 30 fps -> meshes at 60 fps render times
    cs, ms, cr, mr = W.random_pca(len(F), seed=1)
    rec.set_pca(cs, ms, cr, mr)
    xs, xr = W.random_coeffs(20, seed=9)
    t30 = np.arange(20) / 30.0
    t60 = np.arange(39) / 60.0
    xs_q = D.seek_batch(t60, t30, torch.from_numpy(xs).cuda())
    xr_q = D.seek_batch(t60, t30, torch.from_numpy(xr).cuda())
    out = rec.decode_and_get_mesh(xs_q, xr_q).cpu().numpy()
    dg = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float32)  # [20, 89784]: what the reference interpolates
    for i in (0, 1, 18, 37, 38):
        ref = chk.get_mesh(np.asarray(seek(t60[i], t30, dg)).astype(np.float32).astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(out[i] - ref).max() <= tol, i
