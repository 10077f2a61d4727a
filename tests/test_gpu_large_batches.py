"""GPU tier, part 2: parity at the batch sizes the bench times.

The persistent kernels walk several tiles per CTA only above ~19 000 frames (K3T: 148 CTAs x 128 columns / 3
coordinates; K1: one CTA pair per two 128-frame tiles), so the tile loops -- mbarrier phase carry-over between tiles,
ring phases, event parity -- are only exercised by batches of that size.  The trick that makes this cheap and
exhaustive: the batch repeats 128 base frames, every solve / decode / assembly tile therefore holds the same
columns, and EVERY tile of the output must be bit-identical to tile 0 (no atomics, fixed summation orders); tile 0
itself is checked against the compiled reference (oracle/_ref) on sampled frames.
"""
import threading

import numpy as np
import pytest

import deformation as D
from deformation import workloads as W
from oracle import ref_loader
from oracle.dgrad_oracle import TriangleDeformationOracle, pca_decode

pytestmark = pytest.mark.gpu

BASE = 128


def _checker(V, F, cnsts=()):
    o = ref_loader.RefSolver(1) if ref_loader.ref_available() else TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=cnsts)
    return o


@pytest.fixture(scope="module")
def chk(flame):
    return _checker(flame["V"], flame["F"], flame["nfv"])


@pytest.fixture(scope="module")
def pca(flame):
    return W.random_pca(len(flame["F"]), seed=1, zero_tris=flame["nft"])


def _assert_tiles_equal_tile0(out, n, what):
    """out [n, rows, 3] on the device: every BASE-frame block equals the first one (ragged tail included)."""
    import torch
    full = n // BASE
    blocks = out[: full * BASE].view(full, BASE, -1)
    same = (blocks == blocks[0:1]).all(dim=2).all(dim=1)
    bad = torch.nonzero(~same).flatten()
    assert bad.numel() == 0, f"{what}: {bad.numel()} of {full} tiles differ from tile 0, first at tile {int(bad[0])}"
    if n % BASE:
        assert torch.equal(out[full * BASE:], out[: n % BASE]), f"{what}: ragged tail differs"


def _sampled_vs_reference(out0, dg_base, chk, V, nfv, tol, what, frames=(0, 1, 31, 32, 63, 64, 96, 127)):
    for i in frames:
        ref = chk.get_mesh(dg_base[i].astype(np.float64), vert_cnsts=V[nfv])
        got = out0[i].cpu().numpy()
        assert np.abs(got - ref).max() <= tol, (what, i, float(np.abs(got - ref).max()))
        assert np.array_equal(got[nfv], V[nfv])


@pytest.mark.parametrize("solver", ["tensor", "simt"])
def test_20k_frames_every_tile_bit_identical_and_vs_reference(flame, chk, pca, solver):
    """20 096 frames (157 tiles of 128 = 471 K3T tiles over 148 CTAs: 3-4 tiles per CTA; 628 SIMT tiles) through
    both public batched calls."""
    import torch
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    n = 157 * BASE + 37                                   # ragged last tile
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver=solver)
    # --- get_mesh_batch (gather assembly -> solve -> output)
    dg_base = W.iid_dgrad(BASE, len(F), sigma=0.05, seed=41)
    dg = torch.from_numpy(dg_base).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
    out = rec.get_mesh_batch(dg)
    torch.cuda.synchronize()
    _assert_tiles_equal_tile0(out, n, f"get_mesh_batch/{solver}")
    _sampled_vs_reference(out[:BASE], dg_base, chk, V, nfv, tol, f"get_mesh_batch/{solver}")
    del dg, out
    # --- decode_and_get_mesh (K1 -> staged assembly -> solve -> output)
    rec.set_pca(*pca)
    xs_b, xr_b = W.random_coeffs(BASE, seed=43)
    xs = torch.from_numpy(xs_b).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
    xr = torch.from_numpy(xr_b).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
    out = rec.decode_and_get_mesh(xs, xr)
    torch.cuda.synchronize()
    _assert_tiles_equal_tile0(out, n, f"decode_and_get_mesh/{solver}")
    dg_ref = pca_decode(xs_b, pca[0], pca[1], xr_b, pca[2], pca[3], dtype=np.float32)
    _sampled_vs_reference(out[:BASE], dg_ref, chk, V, nfv, tol, f"decode_and_get_mesh/{solver}")
    rec.close()


def test_bench_batch_75600_and_chunked_80000(flame, chk, pca):
    """The exact batch bench.py times (315 sentences x 240 = 75 600 frames: 1773 K3T tiles, 12 per CTA; K1 walks four
    frame-tile pairs per cluster) and one above the chunking threshold (148 x 128 x 4 = 75 776 frames), both through
    decode_and_get_mesh; plus the dgrad-resident call at 75 600 through a strided view of a small buffer."""
    import torch
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
    assert rec.debug("ts_stats")[0] == 1
    rec.set_pca(*pca)
    xs_b, xr_b = W.random_coeffs(BASE, seed=47)
    dg_ref = pca_decode(xs_b, pca[0], pca[1], xr_b, pca[2], pca[3], dtype=np.float32)
    for n in (75600, 80000):
        xs = torch.from_numpy(xs_b).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
        xr = torch.from_numpy(xr_b).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
        out = rec.decode_and_get_mesh(xs, xr)
        torch.cuda.synchronize()
        _assert_tiles_equal_tile0(out, n, f"decode_and_get_mesh/{n}")
        _sampled_vs_reference(out[:BASE], dg_ref, chk, V, nfv, tol, f"decode_and_get_mesh/{n}")
        # the free-rows variant is the same numbers without the constant rows
        free = rec.decode_and_get_mesh(xs, xr, free_only=True)
        ids = torch.from_numpy(rec.free_vertices).cuda().long()
        assert free.shape == (n, rec.n_free, 3)
        assert torch.equal(free[:BASE], out[:BASE][:, ids])
        _assert_tiles_equal_tile0(free, n, f"free_only/{n}")
        assert torch.equal(rec.expand_free(free[-300:]), out[-300:])
        del out, free
    # dgrad-resident path at the bench size: 75 600 rows of a [128, 89784] buffer cannot be expressed as one strided
    # tensor, so materialise it (27 GB) only if the device has the room, else 30 000 frames (704 K3T tiles: 4-5 per CTA)
    free_b, _ = torch.cuda.mem_get_info()
    n = 75600 if free_b > 60 * 2 ** 30 else 30000
    dg = torch.from_numpy(dg_ref).cuda().repeat(n // BASE + 1, 1)[:n].contiguous()
    out = rec.get_mesh_batch(dg)
    torch.cuda.synchronize()
    _assert_tiles_equal_tile0(out, n, f"get_mesh_batch/{n}")
    _sampled_vs_reference(out[:BASE], dg_ref, chk, V, nfv, tol, f"get_mesh_batch/{n}")
    rec.close()


@pytest.mark.parametrize("sigma", [0.8, 2.0])
def test_rotation_angles_above_one_radian(flame, chk, sigma):
    """The theta > 1 branch (csrc/kernels.cu eq_vectors / rot_coeffs: sinf instead of the series) against the
    reference's exp (rotation/utils_rotation.cpp:20-51, always sin/cos): iid dgrad with sigma 0.8 (two thirds of the
    triangles above one radian) and 2.0 (97 %), through the gather assembly, both solvers."""
    V, F, nfv, tol = flame["V"], flame["F"], flame["nfv"], flame["tol"]
    dg = W.iid_dgrad(70, len(F), sigma=sigma, seed=77)
    th = np.sqrt((dg.reshape(70, -1, 9)[..., 6:] ** 2).sum(-1))
    assert (th > 1).mean() > 0.6
    for solver in ("tensor", "simt"):
        rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver=solver)
        out = rec.get_mesh_batch(dg)
        for i in (0, 1, 33, 69):
            ref = chk.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
            assert np.abs(out[i] - ref).max() <= tol, (solver, sigma, i, float(np.abs(out[i] - ref).max()))
        rec.close()


def test_rotation_angles_above_one_radian_staged_path(flame, chk):
    """Same branch in the STAGED assembly (k_assemble, packed two-frames-per-lane arithmetic: rot_coeffs), which only
    the decode path reaches.  The basis is made of a few dyadic entries per row and the coefficients are multiples of
    1/8, so the decode is exact in fp32, in 3xTF32 and in fp64 alike and the comparison isolates the transform."""
    import torch
    V, F, nfv, nft, tol = flame["V"], flame["F"], flame["nfv"], flame["nft"], flame["tol"]
    nt = len(F)
    rng = np.random.default_rng(5)
    ks, kr = 85, 180

    def dyadic_basis(rows, k, nnz, vals):
        Wm = np.zeros((rows, k), dtype=np.float32)
        cols = rng.integers(0, k, size=(rows, nnz))
        Wm[np.arange(rows)[:, None], cols] = rng.choice(vals, size=(rows, nnz)).astype(np.float32)
        return Wm

    cs = dyadic_basis(nt * 6, ks, 2, [0.03125, -0.03125, 0.0625])
    cr = dyadic_basis(nt * 3, kr, 4, [0.5, -0.5, 0.25, -0.25])
    ms = np.zeros(nt * 6, dtype=np.float32)
    mr = (rng.integers(-2, 3, nt * 3) / 8.0).astype(np.float32)
    xs = (np.round(rng.standard_normal((200, ks)) * 8) / 8).astype(np.float32)
    xr = (np.round(rng.standard_normal((200, kr)) * 8) / 8).astype(np.float32)
    dg64 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64)
    dg32 = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float32)
    assert np.array_equal(dg32.astype(np.float64), dg64)            # the decode is exact by construction
    active = np.setdiff1d(np.arange(nt), nft)
    th = np.sqrt((dg32.reshape(200, -1, 9)[:, active, 6:] ** 2).sum(-1))
    assert (th > 1).mean() > 0.3 and (th < 1).mean() > 0.1           # both branches, mixed inside warps
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
    rec.set_pca(cs, ms, cr, mr)
    got = rec.decode_compact(torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()).cpu().numpy()
    lay = rec.compact_layout()
    assert np.array_equal(got[:, lay >= 0], dg32[:, lay[lay >= 0]])   # K1 reproduces it bit for bit
    out = rec.decode_and_get_mesh(xs, xr)
    for i in (0, 63, 64, 65, 127, 128, 199):
        ref = chk.get_mesh(dg64[i], vert_cnsts=V[nfv])
        assert np.abs(out[i] - ref).max() <= tol, (i, float(np.abs(out[i] - ref).max()))
    rec.close()


def test_decode_dgrad_with_correspondences(golden_small):
    """the full-layout decode is sized by the SOURCE triangle count of the basis (11 here), not by the
    template's (96), and refuses to run after the source count changed."""
    import torch
    V, F, border = W.grid_mesh()
    cc, cf = golden_small["corr_count"], golden_small["corr_faces"]
    n_src = 11
    r = D.Reconstructor(V, F, cnsts=border, corrs=cc, device=0)
    r.set_correspondences(cc, cf, n_src_tris=n_src)
    cs, ms, cr, mr = W.random_pca(n_src, seed=8, k_scale=7, k_rotat=9, target_std=0.05)
    r.set_pca(cs, ms, cr, mr)
    xs, xr = W.random_coeffs(33, seed=3, k_scale=7, k_rotat=9)
    xs_d, xr_d = torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()
    dg = r.decode_dgrad(xs_d, xr_d)
    assert dg.shape == (33, n_src * 9)
    ref = pca_decode(xs, cs, ms, xr, cr, mr, dtype=np.float64)
    assert np.abs(dg.cpu().numpy() - ref).max() <= 1e-6
    # decode + reconstruct through the correspondences == reconstructing the decoded source dgrad
    out = r.decode_and_get_mesh(xs_d, xr_d)
    ref_out = r.get_mesh_batch(dg)
    assert float((out - ref_out).abs().max()) <= 0.25e-6 * W.bbox_diag(V)
    # a second basis replaces the first (and frees it); results follow the new one
    cs2, ms2, cr2, mr2 = W.random_pca(n_src, seed=9, k_scale=7, k_rotat=9, target_std=0.05)
    r.set_pca(cs2, ms2, cr2, mr2)
    ref2 = pca_decode(xs, cs2, ms2, xr, cr2, mr2, dtype=np.float64)
    assert np.abs(r.decode_dgrad(xs_d, xr_d).cpu().numpy() - ref2).max() <= 1e-6
    # switching back to "block k reads source triangle k" changes the source count: the basis no longer fits
    r.set_correspondences()
    with pytest.raises(D.SdfaError):
        r.decode_dgrad(xs_d, xr_d)
    r.close()


def test_set_pca_does_not_leak(flame, pca):
    """repeated sdfa_set_pca calls used to keep every previous basis (~60 MB each) until destroy."""
    import torch
    rec = D.Reconstructor(flame["V"], flame["F"], cnsts=flame["nfv"], device=0)
    rec.set_pca(*pca)
    torch.cuda.synchronize()
    before, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        rec.set_pca(*pca)
    torch.cuda.synchronize()
    after, _ = torch.cuda.mem_get_info()
    assert before - after < 32 * 2 ** 20, (before - after) / 2 ** 20
    rec.close()


def test_free_rows_host_and_device(flame, pca):
    """The opt-in free-rows output: same numbers as the reference-layout call, a quarter of the bytes."""
    import torch
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
    ids = rec.free_vertices
    assert np.array_equal(ids, np.setdiff1d(np.arange(len(V)), nfv))
    dg = W.iid_dgrad(70, len(F), sigma=0.02, seed=12)
    full = rec.get_mesh_batch(dg)
    assert np.array_equal(rec.get_mesh_batch(dg, free_only=True), full[:, ids])
    rec.set_pca(*pca)
    xs, xr = W.random_coeffs(5000, seed=13)                 # two host chunks
    full = rec.decode_and_get_mesh(xs, xr)
    free = rec.decode_and_get_mesh(xs, xr, free_only=True)
    assert free.shape == (5000, rec.n_free, 3) and np.array_equal(free, full[:, ids])
    back = rec.expand_free(torch.from_numpy(free[:100]).cuda()).cpu().numpy()
    assert np.array_equal(back, full[:100])
    # moved constraints show up in the expansion's constant rows
    C2 = (V[nfv] + np.float32(0.001)).astype(np.float32)
    rec.set_constraint_positions(C2)
    back = rec.expand_free(rec.decode_and_get_mesh(torch.from_numpy(xs[:10]).cuda(), torch.from_numpy(xr[:10]).cuda(),
                                                   free_only=True)).cpu().numpy()
    assert np.array_equal(back[:, nfv], np.broadcast_to(C2, (10,) + C2.shape))
    assert np.array_equal(back, rec.decode_and_get_mesh(xs[:10], xr[:10]))
    rec.close()


def test_handle_shared_by_threads_and_streams(flame, pca):
    """SURVEY 8(b) "Threading": one handle, four host threads, each on its own CUDA stream, interleaving both batched
    entry points -- every result equals the single-threaded one bit for bit."""
    import torch
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
    rec.set_pca(*pca)
    dg = torch.from_numpy(W.iid_dgrad(300, len(F), sigma=0.03, seed=15)).cuda()
    xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(700, seed=16))
    ref_a, ref_b = rec.get_mesh_batch(dg).clone(), rec.decode_and_get_mesh(xs, xr).clone()
    torch.cuda.synchronize()
    errors = []

    def work(k):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for it in range(6):
                    if (it + k) & 1:
                        out = rec.get_mesh_batch(dg, stream=st.cuda_stream)
                        st.synchronize()
                        assert torch.equal(out, ref_a), (k, it, "get_mesh_batch")
                    else:
                        out = rec.decode_and_get_mesh(xs, xr, stream=st.cuda_stream)
                        st.synchronize()
                        assert torch.equal(out, ref_b), (k, it, "decode_and_get_mesh")
        except Exception as e:                         # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    rec.close()


def test_config5_vs_compiled_reference():
    """Config 5 against the compiled reference itself, several SIMT tiles per CTA included:
    600 frames = 38 tiles of 16 -- and repeated frames must come out identical wherever they sit."""
    import torch
    if not ref_loader.ref_available():
        pytest.skip("oracle/_ref did not travel with the snapshot")
    V, F, c = W.flame_sub2()
    tol = 1e-6 * W.bbox_diag(V)
    r = D.Reconstructor(V, F, cnsts=c, device=0)
    o = ref_loader.RefSolver(1)
    assert o.set_target(V, F, cnsts=c)
    base = W.iid_dgrad(16, len(F), sigma=0.01, seed=5)
    n = 16 * 37 + 8
    dg = torch.from_numpy(base).cuda().repeat(38, 1)[:n].contiguous()
    out = r.get_mesh_batch(dg)
    blocks = out[: 16 * 37].view(37, 16, -1)
    assert bool((blocks == blocks[0:1]).all())
    assert torch.equal(out[16 * 37:], out[:8])
    for i in (0, 7, 15):
        ref = o.get_mesh(base[i].astype(np.float64), vert_cnsts=V[c])
        got = out[i].cpu().numpy()
        assert np.abs(got - ref).max() <= tol, (i, float(np.abs(got - ref).max()))
        assert np.array_equal(got[c], V[c])
    r.close()


def test_kernel_generations_agree(flame):
    """Every build-in alternative of a kernel (options of sdfa_create_with) against the defaults on one batch: the output
    kernels must agree bit for bit whatever their generation and frames per CTA (they move the same values), the two
    assembly generations and the two decode generations within the arithmetic's tolerance, and all of them with the
    compiled reference on sampled frames."""
    import torch
    V, F, nfv, nft, tol = flame["V"], flame["F"], flame["nfv"], flame["nft"], flame["tol"]
    pca = W.random_pca(len(F), seed=1, zero_tris=nft)
    n = 1500                                                   # 23.4 tiles of 64 frames: partial tiles everywhere
    xs, xr = W.random_coeffs(n, seed=21)
    xs_d, xr_d = torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()

    def run(options):
        r = D.Reconstructor(V, F, cnsts=nfv, device=0, options=options)
        r.set_pca(*pca)
        dec = r.decode_and_get_mesh(xs_d, xr_d).cpu().numpy()
        free = r.decode_and_get_mesh(xs_d, xr_d, free_only=True).cpu().numpy()
        dg = r.decode_dgrad(xs_d, xr_d)
        rec = r.get_mesh_batch(dg).cpu().numpy()
        ids = r.free_vertices
        r.close()
        assert np.array_equal(dec[:, ids], free)
        return dec, rec

    base_dec, base_rec = run({})
    for opts in ({"output": 1}, {"output": 2, "output_frames": 32}, {"output": 2, "output_frames": 16}):
        dec, rec = run(opts)
        assert np.array_equal(dec, base_dec) and np.array_equal(rec, base_rec), opts
    dec, rec = run({"asm_gather": 1})
    assert np.array_equal(dec, base_dec) and np.abs(rec - base_rec).max() <= 1e-7
    for gen in (2, 3):                                         # the same arithmetic fed by cp.async rings / by tensor-map boxes: bit-equal
        dec, rec = run({"asm_gather": gen})
        assert np.array_equal(dec, base_dec) and np.array_equal(rec, base_rec), gen
    dec, rec = run({"decode": "tf32"})
    assert np.array_equal(rec, base_rec) and np.abs(dec - base_dec).max() <= 1e-7
    chk = _checker(V, F, nfv)
    dgh = pca_decode(xs, *pca[:2], xr, *pca[2:], dtype=np.float32)
    for i in (0, 63, 64, n - 1):
        want = chk.get_mesh(dgh[i].astype(np.float64), vert_cnsts=V[nfv])
        assert np.abs(base_dec[i] - want).max() <= tol and np.abs(base_rec[i] - want).max() <= tol


def test_no_write_outside_the_callers_buffers(flame):
    """compute-sanitizer is closed on this pool, so the bounds of what the kernels write into CALLER-provided device buffers
    are checked here: every output buffer sits between two guard regions of a recognisable bit pattern that must survive
    the call -- vertices (full and free-rows layout, both entry points), the compact dgrad of sdfa_decode_compact_dev
    (written by the decode kernel in 64-frame tiles, both kernel generations) -- at frame counts around every tile size."""
    import torch
    from deformation._native import check, lib, ptr
    V, F, nfv, nft = flame["V"], flame["F"], flame["nfv"], flame["nft"]
    pca = W.random_pca(len(F), seed=1, zero_tris=nft)
    GUARD, PAT = 1 << 16, 0x7FC0BEEF                           # floats per guard region; a NaN payload no kernel produces

    def guarded(n_floats):
        buf = torch.full((GUARD + n_floats + GUARD,), PAT, dtype=torch.int32, device="cuda")
        return buf, buf[GUARD:GUARD + n_floats].view(torch.float32)

    def intact(buf, n_floats):
        return bool((buf[:GUARD] == PAT).all()) and bool((buf[GUARD + n_floats:] == PAT).all())

    for opts in ({}, {"decode": "tf32"}):
        r = D.Reconstructor(V, F, cnsts=nfv, device=0, options=opts)
        r.set_pca(*pca)
        slots = lib.sdfa_compact_layout(r._h, None, 0)
        T = int(r.debug("compact_tile")[0])
        for n in (1, 31, 33, 63, 64, 65, 127, 128, 129, 255, 257, 300):
            xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(n, seed=n))
            s = torch.cuda.current_stream().cuda_stream
            # vertices, both layouts, decode path
            for free_only, rows in ((False, len(V)), (True, r.n_free)):
                buf, out = guarded(n * rows * 3)
                r.decode_and_get_mesh(xs, xr, out=out.view(n, rows, 3), free_only=free_only)
                torch.cuda.synchronize()
                assert intact(buf, n * rows * 3), (opts, n, free_only)
                assert not bool((out.view(torch.int32) == PAT).any()), (opts, n, free_only)      # and every element was written
            # the compact dgrad: whole 64-frame tiles
            nc = (n + T - 1) // T * slots * T
            buf, cd = guarded(nc)
            check(lib.sdfa_decode_compact_dev(r._h, ptr(xs.data_ptr()), ptr(xr.data_ptr()), n, ptr(cd.data_ptr()), ptr(s)))
            torch.cuda.synchronize()
            assert intact(buf, nc), (opts, n, "compact")
            if opts:
                continue
            # vertices from a reference-layout dgrad
            dg = r.decode_dgrad(xs, xr)
            for free_only, rows in ((False, len(V)), (True, r.n_free)):
                buf, out = guarded(n * rows * 3)
                r.get_mesh_batch(dg, out=out.view(n, rows, 3), free_only=free_only)
                torch.cuda.synchronize()
                assert intact(buf, n * rows * 3), (n, free_only, "dgrad")
                assert not bool((out.view(torch.int32) == PAT).any()), (n, free_only, "dgrad")
        r.close()
