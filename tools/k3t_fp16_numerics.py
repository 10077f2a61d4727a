"""CPU experiment behind DESIGN.md section 4 (K3T, "an FP16 hi/lo split ... numerics checked"): the tensor plan's own op
order (MMA and EPI streams fetched through sdfa_debug_get, run sequentially) with the operands as FP16 hi/lo pairs instead
of TF32 hi/lo pairs -- per-column power-of-two scales for the A operands (one for the forward sweep from the column's
right-hand-side maximum, one for the backward sweep from x_root), one global scale for the matrices, products
a_hi b_hi + a_lo b_hi + a_hi b_lo in fp32, results unscaled when read back -- against the reference on FLAME.
Prints the maximum vertex error per noise level for one scale per column and for the two-scale scheme.
Not part of the product or of the tests; needs the library built (device -1: host plans only) and oracle/.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import deformation as D                                   # noqa: E402
from deformation import workloads as W                    # noqa: E402
from oracle.dgrad_oracle import TriangleDeformationOracle  # noqa: E402
from tests import plan_emulator as E                      # noqa: E402
from tests import tplan_emulator as T                     # noqa: E402

f32, f16 = np.float32, np.float16


def split16(x):
    hi = x.astype(f16)
    lo = (x - hi.astype(f32)).astype(f16)
    return hi.astype(f32), lo.astype(f32)


def solve_tile_fp16(pl, scratch, s_b, fwd_target, bwd_target):
    """scratch [n_rows, columns] f32 in place.  bwd_target None: one scale per column for both sweeps."""
    mma, epi, matrix, chunk_off = pl["mma"], pl["epi"], pl["matrix"], pl["chunk_off"]
    cols = scratch.shape[1]
    r = np.abs(scratch).max(0)
    r[r == 0] = 1
    s_a = (2.0 ** np.floor(np.log2(fwd_target / r))).astype(f32)
    inv = (1.0 / (s_a * f32(s_b))).astype(f32)
    tm = np.full((cols, T.TMEM_COLS), np.nan, f32)
    epi_evt, mma_evt = np.zeros(int(pl["stats"][7]), bool), np.zeros(int(pl["stats"][6]), bool)
    state = {"pm": 0, "chunk": -1, "stage": None, "max_operand": 0.0}

    def run_mma_ready():
        while state["pm"] < len(mma):
            op = mma[state["pm"]]
            if any(w >= 0 and not epi_evt[w] for w in (op["wait_epi"], op["wait_epi2"])):
                break
            if op["flags"] & T.MMA_CHUNK_FIRST:
                state["chunk"] += 1
                state["stage"] = matrix[chunk_off[state["chunk"]]:chunk_off[state["chunk"] + 1]]
            nn, k8 = int(op["n"]), int(op["k8"])
            b = (T._tile(state["stage"], int(op["b_hi_off"]), nn, k8).astype(np.float64) +
                 T._tile(state["stage"], int(op["b_lo_off"]), nn, k8).astype(np.float64))
            b1, b2 = split16((b * s_b).astype(f32))
            d, ah, al = int(op["d_col"]), int(op["a_hi_col"]), int(op["a_lo_col"])
            a1, a2 = tm[:, ah:ah + 8 * k8], tm[:, al:al + 8 * k8]
            state["max_operand"] = max(state["max_operand"], float(np.abs(a1).max()))
            prod = (a1 @ b1.T + a2 @ b1.T + a1 @ b2.T).astype(f32)
            if op["flags"] & T.MMA_ACCUMULATE:
                tm[:, d:d + nn] += prod
            else:
                tm[:, d:d + nn] = prod
            if op["commit_mma"] >= 0:
                mma_evt[op["commit_mma"]] = True
            state["pm"] += 1

    for op in epi:
        run_mma_ready()
        assert op["wait_mma"] < 0 or mma_evt[op["wait_mma"]]
        nch, nv, fl = int(op["n_chunks"]), int(op["n_valid"]), int(op["flags"])
        v = np.zeros((cols, 8 * nch), f32)
        if fl & T.EPI_FROM_TMEM:
            src = int(op["src_col"])
            v[:, :nv] = tm[:, src:src + nv] * inv[:, None]
            if fl & T.EPI_ZERO_SRC:
                tm[:, src:src + 8 * nch] = 0
        if fl & T.EPI_ADD_GLOBAL:
            v[:, :nv] = v[:, :nv] + scratch[int(op["row_in"]):int(op["row_in"]) + nv].T
        v[:, nv:] = 0
        if fl & T.EPI_STORE_GLOBAL:
            scratch[int(op["row_out"]):int(op["row_out"]) + nv] = v[:, :nv].T
        if (fl & T.EPI_AFTER_STORES) and bwd_target is not None:       # x_root: the backward sweep's operand scale
            rx = np.abs(v).max(1)
            rx[rx == 0] = 1
            s_a = (2.0 ** np.floor(np.log2(bwd_target / rx))).astype(f32)
            inv = (1.0 / (s_a * f32(s_b))).astype(f32)
        if fl & T.EPI_ST_RAW:
            tm[:, int(op["hi_col"]):int(op["hi_col"]) + 8 * nch] = v
        if fl & T.EPI_ST_SPLIT:
            vs = (v * s_a[:, None]).astype(f32)
            assert np.abs(vs).max() < 65504, "FP16 overflow"
            hi, lo = split16(vs)
            tm[:, int(op["hi_col"]):int(op["hi_col"]) + 8 * nch] = hi
            tm[:, int(op["lo_col"]):int(op["lo_col"]) + 8 * nch] = lo
        for key in ("signal_epi", "signal_read"):
            if op[key] >= 0:
                epi_evt[op[key]] = True
    run_mma_ready()
    return state["max_operand"]


def main():
    V, F, nfv, _ = W.load_flame()
    rec = D.Reconstructor(V, F, cnsts=nfv, device=-1, solver="tensor")
    pl = T.plan(rec)
    mx, ch, stage = 0.0, -1, None
    for op in pl["mma"]:
        if op["flags"] & T.MMA_CHUNK_FIRST:
            ch += 1
            stage = pl["matrix"][pl["chunk_off"][ch]:pl["chunk_off"][ch + 1]]
        mx = max(mx, float(np.abs(T._tile(stage, int(op["b_hi_off"]), int(op["n"]), int(op["k8"]))).max()))
    s_b = 2.0 ** np.floor(np.log2(2.0 ** 14 / mx))
    o = TriangleDeformationOracle()
    assert o.set_target(V, F, cnsts=nfv)
    tol = 1e-6 * W.bbox_diag(V)
    rows = rec.debug("scratch_row")
    free_to_vi, perm = rec.debug("free_to_vi"), rec.debug("perm")
    iperm = np.empty_like(perm)
    iperm[perm] = np.arange(len(perm))
    xb = rec.debug("x_base").reshape(-1, 3)
    xb_row = np.empty_like(xb)
    xb_row[rows] = xb[iperm]
    vert_of_row = np.empty(rec.n_free, dtype=np.int64)
    vert_of_row[rows] = free_to_vi
    print(f"largest matrix entry {mx:.4f}, matrix scale 2^{int(np.log2(s_b))}, tolerance {tol:.3e} m")
    for sigma in (0.01, 0.2, 2.0):
        dg = W.iid_dgrad(4, len(F), sigma=sigma, seed=1)
        rhs = E.assemble(rec, dg)
        refs = [o.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])[vert_of_row] for i in range(4)]
        for name, bwd in (("one scale per column", None), ("forward + backward scale", 2.0 ** 4)):
            sc = np.zeros((rec.n_free, 128), f32)
            for c in range(3):
                sc[:, c * 4:(c + 1) * 4] = rhs[:, :, c].T
            biggest = solve_tile_fp16(pl, sc, s_b, 2.0 ** 2, bwd)
            worst = 0.0
            for i in range(4):
                out = np.stack([xb_row[:, c] + sc[:, c * 4 + i] for c in range(3)], 1)
                worst = max(worst, float(np.abs(out - refs[i]).max()))
            print(f"sigma {sigma:5.2f}  {name:26s} max |dv| {worst:.3e} m   largest scaled operand {biggest:.1f}")


if __name__ == "__main__":
    main()
