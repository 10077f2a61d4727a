"""CPU tier: the C ABI loads and exports what include/sdfa_b200.h declares, and the host-side analysis
(system matrix, ordering, Cholesky factor, solve program, assembly plan) is right -- checked by running
the uploaded plans through the test-only numpy interpreter (tests/plan_emulator.py) against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import deformation as D
from deformation import _native, workloads as W
from oracle.dgrad_oracle import TriangleDeformationOracle
from tests import plan_emulator as E
from tests import tplan_emulator as T

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sdfa_b200.h")).read()
    declared = set(re.findall(r"\b(sdfa_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("sdfa_handle")
    assert declared == set(_native.EXPORTS)
    lib = ctypes.CDLL(os.path.abspath(_native.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name


def test_no_device_means_error_not_fallback(flame):
    r = D.Reconstructor(flame["V"], flame["F"], cnsts=flame["nfv"], device=-1)
    with pytest.raises(D.SdfaError) as ei:
        r.get_mesh_batch(np.zeros((1, 9976 * 9), dtype=np.float32))
    assert ei.value.code == _native.ERR_CUDA
    with pytest.raises(D.SdfaError):
        r.get_mesh(np.zeros(9976 * 9), vert_cnsts=flame["V"][flame["nfv"]])


def test_argument_errors_raise(flame):
    V, F = flame["V"], flame["F"]
    with pytest.raises(D.SdfaError):
        D.Reconstructor(V, F, cnsts=[0, 0], device=-1)             # duplicate constraint
    with pytest.raises(D.SdfaError):
        D.Reconstructor(V, F, cnsts=[len(V)], device=-1)           # out of range
    bad = F.copy(); bad[0, 0] = len(V)
    with pytest.raises(D.SdfaError):
        D.Reconstructor(V, bad, device=-1)
    with pytest.raises(D.SdfaError):
        D.Reconstructor(V, F, cnsts=np.arange(len(V)), device=-1)  # nothing left to solve


@pytest.fixture(scope="module")
def flame_rec(flame):
    """SIMT solve plan (the sweeps of kernel K3); the tensor-core plan has its own fixture below."""
    return D.Reconstructor(flame["V"], flame["F"], cnsts=flame["nfv"], device=-1, solver="simt")


@pytest.fixture(scope="module")
def flame_rec_tensor(flame):
    return D.Reconstructor(flame["V"], flame["F"], cnsts=flame["nfv"], device=-1)


@pytest.fixture(scope="module")
def flame_oracle(flame):
    o = TriangleDeformationOracle()
    assert o.set_target(flame["V"], flame["F"], cnsts=flame["nfv"])
    return o


def _csc(colptr, rowidx, val, n):
    return sp.csc_matrix((val, rowidx, colptr), shape=(n, n))


def test_system_matrix_and_factor(flame_rec, flame_oracle):
    r, o = flame_rec, flame_oracle
    n = r.n_free
    assert (r.n_free, r.n_eq, r.n_active) == (1261, 9976, 2601)
    Ml = _csc(r.debug("m_colptr"), r.debug("m_rowidx"), r.debug("m_val"), n)
    M = Ml + sp.tril(Ml, -1).T
    assert abs(M - o.AtA).max() < 1e-9 * abs(o.AtA).max()
    perm = r.debug("perm")
    assert sorted(perm) == list(range(n))
    L = _csc(r.debug("l_colptr"), r.debug("l_rowidx"), r.debug("l_val"), n)
    assert sp.triu(L, 1).nnz == 0
    PMP = M.tocsr()[perm][:, perm]
    assert abs(L @ L.T - PMP).max() < 1e-10 * abs(PMP).max()
    assert r.nnz_l < 30000                      # fill stays near AMD's 20 385
    parent = r.debug("parent")
    assert all(parent[j] > j or parent[j] < 0 for j in range(n))   # postordered elimination tree


def test_program_budget(flame_rec):
    n_slots, n_phases, st_f, st_b, n_entries, n_stages, nbytes, max_eq, _, _, smem = flame_rec.debug("stats")[:11]
    # work per sweep stays within 2x of the factor's nonzeros (inverted subtree blocks add some fill)
    assert 2 * flame_rec.nnz_l <= n_entries <= 4 * flame_rec.nnz_l
    assert st_f + st_b < 500                  # levels (= consumer barriers) per tile; the column count is 2 x 1261
    assert smem <= 227 * 1024
    assert max_eq * 36 <= 48 * 1024


def test_emulated_flame_frames_match_oracle(flame_rec, flame_oracle, flame, golden_flame):
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    dg = W.iid_dgrad(2, len(F), sigma=0.01, seed=0)
    out, st = E.solve(flame_rec, E.assemble(flame_rec, dg))
    assert st["max_slot_used"] < st["n_slots"]
    assert not np.isnan(out).any()
    for i in range(2):
        assert np.abs(out[i][free] - golden_flame["iid_free_verts"][i]).max() < 0.1 * flame["tol"]
        assert np.array_equal(out[i][nfv], V[nfv])


def test_emulated_large_deformation(flame_rec, flame_oracle, flame, golden_flame):
    V, F, nfv, nft = flame["V"], flame["F"], flame["nfv"], flame["nft"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    active = np.setdiff1d(np.arange(len(F)), nft)
    full = np.zeros((1, len(F), 9), dtype=np.float32)
    full[0, active] = golden_flame["integ30_dgrad_active"]
    out, _ = E.solve(flame_rec, E.assemble(flame_rec, full.reshape(1, -1)))
    assert np.abs(out[0][free] - golden_flame["integ30_free_verts"]).max() < flame["tol"]


def test_emulated_small_mesh_modes(golden_small):
    V, F, border = W.grid_mesh()
    dg = W.iid_dgrad(4, len(F), sigma=0.05, seed=7)
    tol = 1e-6 * W.bbox_diag(V)
    r = D.Reconstructor(V, F, cnsts=border, device=-1, solver="simt")
    out, _ = E.solve(r, E.assemble(r, dg))
    assert np.abs(out - golden_small["cnst_verts"]).max() < tol
    # raw matrices
    dm = golden_small["deform_mat"].astype(np.float32).reshape(1, -1)
    out, _ = E.solve(r, E.assemble(r, dm, mode="matrix"))
    assert np.abs(out[0] - golden_small["from_dm_verts"]).max() < tol
    # moved constraints: the base solution is recomputed on the host
    Cm = (V[border] + np.float32(0.001)).astype(np.float32)
    r.set_constraint_positions(Cm)
    out, _ = E.solve(r, E.assemble(r, dg[:1]), cnst_pos=Cm)
    assert np.abs(out[0] - golden_small["moved_cnst_verts"]).max() < tol
    # single constraint
    r1 = D.Reconstructor(V, F, cnsts=[5], device=-1, solver="simt")
    out, _ = E.solve(r1, E.assemble(r1, dg[1:2]))
    assert np.abs(out[0] - golden_small["one_cnst_verts"]).max() < 5 * tol    # cond ~1e8: fp32 sweeps
    # correspondences
    cc, cf = golden_small["corr_count"], golden_small["corr_faces"]
    rc = D.Reconstructor(V, F, cnsts=border, corrs=cc, device=-1, solver="simt")
    rc.set_correspondences(cc, cf, n_src_tris=11)
    src = W.iid_dgrad(1, 11, sigma=0.05, seed=9)
    out, _ = E.solve(rc, E.assemble(rc, src))
    assert np.abs(out[0] - golden_small["corr_verts"]).max() < tol


def test_emulated_unconstrained_is_pinned(golden_small):
    V, F, _ = W.grid_mesh()
    dg = W.iid_dgrad(4, len(F), sigma=0.05, seed=7)
    r = D.Reconstructor(V, F, device=-1)
    out, _ = E.solve(r, E.assemble(r, dg[2:3]))
    ref = golden_small["uncnst_verts"]
    assert np.abs((out[0] - out[0].mean(0)) - (ref - ref.mean(0))).max() <= 2e-6
    assert np.abs(out[0]).max() < 1.0


def test_emulated_tile_boundaries():
    """33 frames = one full tile + a 1-frame tile; every frame must equal its single-frame result."""
    V, F, border = W.grid_mesh()
    r = D.Reconstructor(V, F, cnsts=border, device=-1, solver="simt")
    dg = W.iid_dgrad(33, len(F), sigma=0.05, seed=3)
    out, _ = E.solve(r, E.assemble(r, dg))
    one, _ = E.solve(r, E.assemble(r, dg[32:33]))
    assert np.array_equal(out[32], one[0])
    o = TriangleDeformationOracle(); o.set_target(V, F, cnsts=border)
    ref = o.get_mesh(dg[17].astype(np.float64), vert_cnsts=V[border])
    assert np.abs(out[17] - ref).max() < 1e-6 * W.bbox_diag(V)


def test_emulated_config5_picks_smaller_tiles():
    """The subdivided template (config 5) does not fit 32 frames per tile; the plan must fall back to 16 or 8
    and still be right (one frame through the interpreter)."""
    V, F, c = W.flame_sub2()
    r = D.Reconstructor(V, F, cnsts=c, device=-1)
    stats = r.debug("stats")
    assert r.n_free == 20653 and stats[14] in (8, 16) and stats[10] <= 227 * 1024
    o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=c)
    dg = W.iid_dgrad(1, len(F), sigma=0.01, seed=5)
    out, _ = E.solve(r, E.assemble(r, dg))
    ref = o.get_mesh(dg[0].astype(np.float64), vert_cnsts=V[c])
    assert np.abs(out[0] - ref).max() <= 0.2e-6 * W.bbox_diag(V)


# ------------------------------------------------------------------------------------------------
# Tensor-core solve plan (csrc/tplan.cpp -> kernel K3T): interpreted by tests/tplan_emulator.py

def test_tensor_plan_budget(flame_rec_tensor):
    st = flame_rec_tensor.debug("ts_stats")
    used, valid, n_mma, n_epi, n_chunks, nbytes, ev_m, ev_e, n_nodes, n_leaves, nk, t_f, t_b, smem, n_streams, n_ring = [int(x) for x in st]
    assert used == 1 and valid == 1 and n_streams == 2
    epi = flame_rec_tensor.debug("ts_epi").view(T.EPI_DT)
    assert n_ring == int(((epi["flags"] & (T.EPI_ADD_GLOBAL | T.EPI_STORE_GLOBAL)) > 0).sum())
    # the forward sweep's stores all run on one stream (its threads fence them for the backward sweep's bulk loads)
    fwd_store = ((epi["flags"] & 7) == 5)
    assert len(set(epi["stream"][fwd_store])) == 1 and (epi["stream"] <= 1).all()
    assert t_f <= 512 and t_b <= 512 and ev_m <= 256 and ev_e <= 256 and smem <= 227 * 1024
    assert nk <= 300_000                      # N*K summed over the products of one 128-column tile
    rows = flame_rec_tensor.debug("scratch_row")
    assert sorted(rows) == list(range(flame_rec_tensor.n_free))
    mma = flame_rec_tensor.debug("ts_mma").view(T.MMA_DT)
    assert (mma["n"] % 16 == 0).all() and (mma["n"] <= 64).all() and (mma["k8"] >= 1).all() and (mma["k8"] <= 8).all()
    assert (mma["d_col"] % 16 == 0).all() and (mma["a_hi_col"] % 8 == 0).all() and (mma["a_lo_col"] % 8 == 0).all()
    assert (mma["b_hi_off"] % 1024 == 0).all() and (mma["b_lo_off"] % 1024 == 0).all()


def test_tensor_emulated_flame_frames_match_oracle(flame_rec_tensor, flame, golden_flame):
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    dg = W.iid_dgrad(2, len(F), sigma=0.01, seed=0)
    rhs = E.assemble(flame_rec_tensor, dg)
    first = None
    for seed in (0, 1, 2, 3):                 # different interleavings of the two instruction streams
        out = T.solve(flame_rec_tensor, rhs, seed=seed)
        assert not np.isnan(out).any()
        for i in range(2):
            assert np.abs(out[i][free] - golden_flame["iid_free_verts"][i]).max() < 0.1 * flame["tol"]
            assert np.array_equal(out[i][nfv], V[nfv])
        if first is None:
            first = out
        assert np.array_equal(out, first)     # the result does not depend on the interleaving


def test_tensor_emulated_plan_variants(flame, golden_flame):
    """One epilogue stream (SDFA_TS_STREAMS=1) and no separate read events (SDFA_TS_EARLY=0): the same results as the
    default plan under the emulator's random interleavings."""
    V, F, nfv = flame["V"], flame["F"], flame["nfv"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    dg = W.iid_dgrad(2, len(F), sigma=0.01, seed=0)
    outs = []
    for env in ({}, {"SDFA_TS_STREAMS": "1"}, {"SDFA_TS_EARLY": "0"}):
        os.environ.update(env)
        try:
            r = D.Reconstructor(V, F, cnsts=nfv, device=-1, solver="tensor")
        finally:
            for k in env:
                del os.environ[k]
        epi = r.debug("ts_epi").view(T.EPI_DT)
        if env.get("SDFA_TS_STREAMS") == "1":
            assert (epi["stream"] == 0).all()
        if env.get("SDFA_TS_EARLY") == "0":
            assert (epi["signal_read"] < 0).all()
        out = T.solve(r, E.assemble(r, dg), seed=5)
        for i in range(2):
            assert np.abs(out[i][free] - golden_flame["iid_free_verts"][i]).max() < 0.1 * flame["tol"]
        outs.append(out)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_tensor_emulated_large_deformation(flame_rec_tensor, flame, golden_flame):
    V, F, nfv, nft = flame["V"], flame["F"], flame["nfv"], flame["nft"]
    free = np.setdiff1d(np.arange(len(V)), nfv)
    active = np.setdiff1d(np.arange(len(F)), nft)
    full = np.zeros((1, len(F), 9), dtype=np.float32)
    full[0, active] = golden_flame["integ30_dgrad_active"]
    out = T.solve(flame_rec_tensor, E.assemble(flame_rec_tensor, full.reshape(1, -1)))
    assert np.abs(out[0][free] - golden_flame["integ30_free_verts"]).max() < 0.25 * flame["tol"]


def test_tensor_emulated_small_mesh_modes(golden_small):
    V, F, border = W.grid_mesh()
    dg = W.iid_dgrad(4, len(F), sigma=0.05, seed=7)
    tol = 1e-6 * W.bbox_diag(V)
    for leaf in ("64", "8"):                  # one leaf, and a real tree on the 35 unknowns
        os.environ["SDFA_TS_LEAF"] = leaf
        try:
            r = D.Reconstructor(V, F, cnsts=border, device=-1, solver="tensor")
        finally:
            del os.environ["SDFA_TS_LEAF"]
        out = T.solve(r, E.assemble(r, dg))
        assert np.abs(out - golden_small["cnst_verts"]).max() < tol
        Cm = (V[border] + np.float32(0.001)).astype(np.float32)
        r.set_constraint_positions(Cm)
        out = T.solve(r, E.assemble(r, dg[:1]), cnst_pos=Cm)
        assert np.abs(out[0] - golden_small["moved_cnst_verts"]).max() < tol
    cc, cf = golden_small["corr_count"], golden_small["corr_faces"]
    rc = D.Reconstructor(V, F, cnsts=border, corrs=cc, device=-1, solver="tensor")
    rc.set_correspondences(cc, cf, n_src_tris=11)
    src = W.iid_dgrad(1, 11, sigma=0.05, seed=9)
    out = T.solve(rc, E.assemble(rc, src))
    assert np.abs(out[0] - golden_small["corr_verts"]).max() < tol


def test_tensor_plan_declines_what_it_cannot_do():
    V, F, _ = W.grid_mesh()
    r = D.Reconstructor(V, F, device=-1)                          # unconstrained: singular system, pinned pivots
    assert r.debug("ts_stats")[0] == 0 and b"pivot" in bytes(r.debug("ts_why_not"))
    with pytest.raises(D.SdfaError):
        D.Reconstructor(V, F, device=-1, solver="tensor")
    V, F, c = W.flame_sub2()
    r = D.Reconstructor(V, F, cnsts=c, device=-1)                 # config 5: 20 653 unknowns -> SIMT sweeps
    assert r.debug("ts_stats")[0] == 0 and b"unknowns" in bytes(r.debug("ts_why_not"))
