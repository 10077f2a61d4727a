"""TEST-ONLY numpy interpreter of the tensor-core solve plan (csrc/tplan.cpp -> kernel K3T, csrc/solve_tc.cu).

Runs the MMA stream and the two EPI instruction streams fetched through ``sdfa_debug_get`` the way the kernel does:
sequential streams that only synchronise through the plan's events, interleaved by a seeded random scheduler,
MMAs completing asynchronously and in order, tensor memory and the matrix ring poisoned with NaN, and the
arithmetic restated as 3xTF32 (operands truncated to TF32, fp32 accumulation).  A missing event shows up as
a NaN / wrong result under some interleaving or as a deadlock.  It is NOT a fallback: nothing under
``sdfa-2019_b200/`` imports it.
"""
import numpy as np

f32 = np.float32
EPI_FROM_TMEM, EPI_ADD_GLOBAL, EPI_STORE_GLOBAL, EPI_ST_RAW, EPI_ST_SPLIT = 1, 2, 4, 8, 16
EPI_LAST_FWD_STORE, EPI_AFTER_STORES, EPI_ZERO_SRC = 32, 64, 128
MMA_ACCUMULATE, MMA_CHUNK_FIRST, MMA_CHUNK_LAST = 1, 2, 4
TS_COLS, STAGE_BYTES, TMEM_COLS = 128, 32768, 512

EPI_DT = np.dtype([("wait_mma", "<i2"), ("signal_epi", "<i2"), ("n_chunks", "<u2"), ("n_valid", "<u2"),
                   ("src_col", "<u2"), ("hi_col", "<u2"), ("lo_col", "<u2"), ("flags", "<u2"),
                   ("row_in", "<u4"), ("row_out", "<u4"), ("stream", "<u2"), ("wait_epi", "<i2"), ("ring_seq", "<u2"), ("signal_read", "<i2")])
MMA_DT = np.dtype([("wait_epi", "<i2"), ("commit_mma", "<i2"), ("d_col", "<u2"), ("a_hi_col", "<u2"),
                   ("a_lo_col", "<u2"), ("n", "<u2"), ("k8", "<u2"), ("flags", "<u2"),
                   ("b_hi_off", "<u4"), ("b_lo_off", "<u4"), ("wait_epi2", "<i2"), ("pad", "<u2"), ("r1", "<u4")])


def tf32(x):
    return (np.ascontiguousarray(x, dtype=f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def _swz(r, k):
    return (r >> 3) * 256 + (r & 7) * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3))


def plan(rec):
    st = rec.debug("ts_stats")
    assert st[1] == 1, "no tensor plan: " + bytes(rec.debug("ts_why_not")).decode()
    return dict(mma=rec.debug("ts_mma").view(MMA_DT), epi=rec.debug("ts_epi").view(EPI_DT),
                matrix=rec.debug("ts_matrix"), chunk_off=rec.debug("ts_chunk_off"), stats=st)


def _tile(stage, off, n, k8):
    """[n, 8*k8] float32 matrix out of the K-major SWIZZLE_128B images of one product (as the UMMA descriptor walks it)."""
    img = stage[off:off + ((k8 + 3) // 4) * n * 128].view(f32)
    rr, kk = np.meshgrid(np.arange(n), np.arange(8 * k8), indexing="ij")
    idx = (kk // 32) * (n * 32) + _swz(rr, kk % 32)
    return img[idx]


def solve_tile(pl, scratch, seed=0, columns=TS_COLS):
    """scratch: [n_rows, columns] f32 (rhs in, x out, in place) for one 128-column tile."""
    rng = np.random.default_rng(seed)
    mma, epi, matrix, chunk_off = pl["mma"], pl["epi"], pl["matrix"], pl["chunk_off"]
    tmem = np.full((columns, TMEM_COLS), np.nan, dtype=f32)
    n_mma_evt, n_epi_evt = int(pl["stats"][6]), int(pl["stats"][7])
    mma_evt, epi_evt = np.zeros(n_mma_evt, bool), np.zeros(n_epi_evt, bool)
    pending = []                          # issued, not yet executed MMA work: ("mma", closure) / ("commit", evt)
    pm = 0
    streams = [np.flatnonzero(epi["stream"] == q) for q in (0, 1)]   # op indices of each EPI stream, in order
    pe = [0, 0]
    held = [None, None]                   # an EPI stream between the two halves of an op: its register values
    assert len(streams[0]) + len(streams[1]) == len(epi)
    ring = epi["ring_seq"][(epi["flags"] & (EPI_ADD_GLOBAL | EPI_STORE_GLOBAL)) > 0]
    assert np.array_equal(ring, np.arange(len(ring))), "ops that use the row ring are numbered in list order"
    chunk = -1
    stage = None

    def run_pending(k):
        for _ in range(k):
            if not pending:
                return
            kind, x = pending.pop(0)
            if kind == "mma":
                x()
            else:
                mma_evt[x] = True

    def issue(op, stage_bytes):
        n, k8 = int(op["n"]), int(op["k8"])
        bh, bl = _tile(stage_bytes, int(op["b_hi_off"]), n, k8), _tile(stage_bytes, int(op["b_lo_off"]), n, k8)
        assert not (np.isnan(bh).any() or np.isnan(bl).any())
        d, ah, al, acc = int(op["d_col"]), int(op["a_hi_col"]), int(op["a_lo_col"]), bool(op["flags"] & MMA_ACCUMULATE)

        def go():
            a_hi, a_lo = tf32(tmem[:, ah:ah + 8 * k8]), tf32(tmem[:, al:al + 8 * k8])
            assert not np.isnan(a_hi).any() and not np.isnan(a_lo).any(), "MMA reads tensor memory that was never written"
            prod = (a_hi @ tf32(bh).T + a_lo @ tf32(bh).T + a_hi @ tf32(bl).T).astype(f32)
            if acc:
                assert not np.isnan(tmem[:, d:d + n]).any(), "MMA accumulates onto unwritten tensor memory"
                tmem[:, d:d + n] += prod
            else:
                tmem[:, d:d + n] = prod
        pending.append(("mma", go))

    def epi_read(op):
        """First half of an EPI op: rows + tensor-memory source -> registers (after it the op signals `signal_read`)."""
        nch, nv, fl = int(op["n_chunks"]), int(op["n_valid"]), int(op["flags"])
        v = np.zeros((columns, 8 * nch), dtype=f32)
        if fl & EPI_ADD_GLOBAL:
            v[:, :nv] = scratch[int(op["row_in"]):int(op["row_in"]) + nv].T
        if fl & EPI_FROM_TMEM:
            t = tmem[:, int(op["src_col"]):int(op["src_col"]) + 8 * nch]
            assert not np.isnan(t[:, :nv]).any(), "EPI reads tensor memory that was never written"
            v[:, :nv] = v[:, :nv] + t[:, :nv]
            if fl & EPI_ZERO_SRC:
                assert op["signal_read"] < 0
                tmem[:, int(op["src_col"]):int(op["src_col"]) + 8 * nch] = 0
        v[:, nv:] = 0
        return v

    def epi_write(op, v):
        """Second half: registers -> scratch rows / tensor memory (after it the op signals `signal_epi`)."""
        nch, nv, fl = int(op["n_chunks"]), int(op["n_valid"]), int(op["flags"])
        if fl & EPI_STORE_GLOBAL:
            scratch[int(op["row_out"]):int(op["row_out"]) + nv] = v[:, :nv].T
        if fl & EPI_ST_RAW:
            tmem[:, int(op["hi_col"]):int(op["hi_col"]) + 8 * nch] = v
        if fl & EPI_ST_SPLIT:
            hi = tf32(v)
            tmem[:, int(op["hi_col"]):int(op["hi_col"]) + 8 * nch] = hi
            tmem[:, int(op["lo_col"]):int(op["lo_col"]) + 8 * nch] = tf32(v - hi)

    # the row loader fetches the rows of the EPI_ADD_GLOBAL ops in list order and holds the first backward op's rows
    # (EPI_AFTER_STORES) back until the forward sweep's last store has been fenced (bar_fwd in the kernel)
    after = np.flatnonzero((epi["flags"] & EPI_AFTER_STORES) > 0)
    first_bwd_seq = int(epi["ring_seq"][after[0]]) if len(after) else 1 << 30
    fwd_done = [False]

    def epi_blocked(q):
        if pe[q] >= len(streams[q]):
            return True
        if held[q] is not None:
            return False
        op = epi[streams[q][pe[q]]]
        if (op["flags"] & EPI_ADD_GLOBAL) and int(op["ring_seq"]) >= first_bwd_seq and not fwd_done[0]:
            return True
        return (op["wait_mma"] >= 0 and not mma_evt[op["wait_mma"]]) or (op["wait_epi"] >= 0 and not epi_evt[op["wait_epi"]])

    def mma_blocked():
        if pm >= len(mma):
            return True
        op = mma[pm]
        return any(w >= 0 and not epi_evt[w] for w in (op["wait_epi"], op["wait_epi2"]))

    steps = 0
    while pe[0] < len(streams[0]) or pe[1] < len(streams[1]) or pm < len(mma) or pending:
        steps += 1
        assert steps < 400000, "deadlock"
        choice = rng.integers(0, 4)
        if choice == 0 and pending:
            run_pending(int(rng.integers(1, 4)))
        elif choice == 1 and pm < len(mma):
            op = mma[pm]
            if mma_blocked():
                if not pending and epi_blocked(0) and epi_blocked(1):
                    raise AssertionError(f"deadlock: mma op {pm}, epi ops {pe} wait for each other")
                continue
            if op["flags"] & MMA_CHUNK_FIRST:
                chunk += 1
                stage = np.full(STAGE_BYTES, 0xFF, dtype=np.uint8)      # NaN-poisoned ring stage
                nb = int(chunk_off[chunk + 1] - chunk_off[chunk])
                assert nb <= STAGE_BYTES and chunk_off[chunk] % 1024 == 0
                stage[:nb] = matrix[chunk_off[chunk]:chunk_off[chunk + 1]]
            issue(op, stage)
            if op["commit_mma"] >= 0:
                pending.append(("commit", int(op["commit_mma"])))
            pm += 1
        elif choice >= 2:
            q = int(choice - 2)
            if pe[q] >= len(streams[q]):
                continue
            if epi_blocked(q):
                run_pending(1)
                continue
            op = epi[streams[q][pe[q]]]
            if held[q] is None:
                held[q] = epi_read(op)
                if op["signal_read"] >= 0:
                    epi_evt[op["signal_read"]] = True
                continue
            epi_write(op, held[q])
            held[q] = None
            if op["flags"] & EPI_LAST_FWD_STORE:
                fwd_done[0] = True
            if op["signal_epi"] >= 0:
                epi_evt[op["signal_epi"]] = True
            pe[q] += 1
    assert chunk + 2 == len(chunk_off)
    return scratch


def solve(rec, rhs, cnst_pos=None, seed=0):
    """K3T + K5: rhs [N, n_free, 3] f32 (scratch-row order) -> verts [N, n_verts, 3] f32."""
    pl = plan(rec)
    N = rhs.shape[0]
    rows = rec.debug("scratch_row")                      # free column -> scratch row
    free_to_vi, perm = rec.debug("free_to_vi"), rec.debug("perm")
    iperm = np.empty_like(perm); iperm[perm] = np.arange(len(perm))
    xb = rec.debug("x_base").reshape(-1, 3)              # Cholesky order: row iperm[f]
    xb_row = np.empty_like(xb); xb_row[rows] = xb[iperm]
    xb_hi = xb_row.astype(f32); xb_lo = (xb_row - xb_hi.astype(np.float64)).astype(f32)
    vert_of_row = np.empty(rec.n_free, dtype=np.int64); vert_of_row[rows] = free_to_vi
    out = np.full((N, rec.n_verts, 3), np.nan, dtype=f32)
    per = TS_COLS
    for f0 in range(0, N, per):
        nf = min(per, N - f0)
        for c in range(3):
            sc = np.zeros((rec.n_free, per), dtype=f32)
            sc[:, :nf] = rhs[f0:f0 + nf, :, c].T
            solve_tile(pl, sc, seed=seed + 7 * c + f0)
            assert not np.isnan(sc).any()
            out[f0:f0 + nf, vert_of_row, c] = (xb_hi[:, c][:, None] + (xb_lo[:, c][:, None] + sc[:, :nf])).T
    if rec.n_cnsts:
        C = rec._verts[rec._cnsts] if cnst_pos is None else np.asarray(cnst_pos, dtype=f32).reshape(-1, 3)
        out[:, rec._cnsts] = C[None]
    return out
