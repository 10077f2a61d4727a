"""TEST-ONLY numpy interpreter of the plans that sdfa_create() uploads to the GPU.

It executes, in float32 and with the kernels' exact data flow, the assembly plan (kernel K2) and the
solve byte-program (kernel K3) fetched through ``sdfa_debug_get`` from a handle built with
``device=-1``.  This lets the CPU test tier validate the host-side scheduler (pieces, slots, levels,
op encoding) without a GPU.  It is NOT a fallback: nothing under ``sdfa-2019_b200/`` imports it.
"""
import numpy as np

OP_ROWS, OP_PHASE_BEGIN, OP_PHASE_END = 1, 2, 3
TASK_GROUP, TASK_OVERWRITE = 1 << 24, 1 << 25
f32 = np.float32


def assemble(rec, dgrad, mode="dgrad"):
    """K2: dgrad [N, n_src*9] f32 -> rhs [N, n_free, 3] f32 (permuted row order)."""
    dg = np.ascontiguousarray(dgrad, dtype=f32).reshape(len(dgrad), -1, 9)
    blocks = rec.debug("asm_blocks").reshape(-1, 5)
    eq_id, eq_u = rec.debug("asm_eq_id"), rec.debug("asm_eq_u").reshape(-1, 8)[:, :6]
    row_perm, eq_rows = rec.debug("asm_row_perm"), rec.debug("asm_eq_rows").reshape(-1, 4)
    colour_ptr = rec.debug("asm_colour_ptr").reshape(len(blocks), -1)
    eq_src = rec.debug("eq_src")
    N = dg.shape[0]
    rhs = np.zeros((N, rec.n_free, 3), dtype=f32)
    for bi, (eb, ee, rb, re, ncol) in enumerate(blocks):
        ids = eq_id[eb:ee]
        src = eq_src[ids]
        u0, u1 = eq_u[eb:ee, :3], eq_u[eb:ee, 3:]
        g = np.zeros((N, ee - eb, 3, 3), dtype=f32)          # [frame][eq][corner][xyz]
        has = src >= 0
        if has.any():
            d = dg[:, src[has]]                               # [N, m, 9]
            if mode == "dgrad":
                th2 = d[..., 6] * d[..., 6] + d[..., 7] * d[..., 7] + d[..., 8] * d[..., 8]
                one = f32(1)
                ts = np.minimum(th2, one)
                a_s = one - ts * f32(1 / 6) * (one - ts * f32(1 / 20) * (one - ts * f32(1 / 42) * (one - ts * f32(1 / 72))))
                b_s = f32(0.5) - ts * f32(1 / 24) * (one - ts * f32(1 / 30) * (one - ts * f32(1 / 56) * (one - ts * f32(1 / 90))))
                thl = np.sqrt(np.maximum(th2, one))
                a_l = np.sin(thl) / thl
                b_l = f32(2) * np.sin(f32(0.5) * thl) ** 2 / np.maximum(th2, one)
                a = np.where(th2 <= one, a_s, a_l)
                b = np.where(th2 <= one, b_s, b_l)
                a = np.where(th2 >= f32(1e-12), a, f32(0)).astype(f32)
                b = np.where(th2 >= f32(1e-12), b, f32(0)).astype(f32)

                def corner(u):
                    u = np.broadcast_to(u[None], d.shape[:2] + (3,))
                    t = np.stack([d[..., 0] * u[..., 0] + d[..., 1] * u[..., 1] + d[..., 2] * u[..., 2],
                                  d[..., 1] * u[..., 0] + d[..., 3] * u[..., 1] + d[..., 4] * u[..., 2],
                                  d[..., 2] * u[..., 0] + d[..., 4] * u[..., 1] + d[..., 5] * u[..., 2]], -1)
                    s = u + t

                    def W(v):
                        return np.stack([d[..., 6] * v[..., 1] + d[..., 7] * v[..., 2],
                                         -d[..., 6] * v[..., 0] + d[..., 8] * v[..., 2],
                                         -d[..., 7] * v[..., 0] - d[..., 8] * v[..., 1]], -1)
                    p = W(s)
                    q = W(p)
                    return (t + a[..., None] * p + b[..., None] * q).astype(f32)
                g2, g3 = corner(u0[has]), corner(u1[has])
            else:
                E = d.reshape(N, -1, 3, 3) - np.eye(3, dtype=f32)
                g2 = np.einsum("nmcr,mr->nmc", E, u0[has]).astype(f32)
                g3 = np.einsum("nmcr,mr->nmc", E, u1[has]).astype(f32)
            g[:, has, 1], g[:, has, 2] = g2, g3
        zero = src == -2
        if zero.any():
            g[:, zero, 1], g[:, zero, 2] = -u0[zero], -u1[zero]
        # the kernel's order: colours in sequence, an equation's corners 0, 1, 2; no row twice within a colour
        acc = np.zeros((N, re - rb, 3), dtype=f32)
        cp = colour_ptr[bi]
        assert cp[0] == 0 and cp[ncol] == ee - eb
        for k in range(ncol):
            touched = set()
            for e in range(cp[k], cp[k + 1]):
                if src[e] == -1:
                    continue
                r0, r1, r2 = (int(x) for x in eq_rows[eb + e][:3])
                mine = {r for r in (r0, r1, r2) if r >= 0}
                assert not (mine & touched), "two equations of one colour share a row"
                touched |= mine
                if r0 >= 0:
                    acc[:, r0] = acc[:, r0] - (g[:, e, 1] + g[:, e, 2])
                if r1 >= 0:
                    acc[:, r1] = acc[:, r1] + g[:, e, 1]
                if r2 >= 0:
                    acc[:, r2] = acc[:, r2] + g[:, e, 2]
        rhs[:, row_perm[rb:re]] = acc
    return rhs


class _Hazards:
    """Slot-granular race detector for the unsynchronised window between two consumer barriers:
    a task may not read what another task wrote, nor write what another task read or wrote."""

    def __init__(self):
        self.sync()

    def sync(self):
        self.r, self.w = {}, {}
        self.n = 0

    def access(self, reads, writes):
        me = self.n
        self.n += 1
        for s in reads:
            assert s not in self.w, f"RAW hazard on slot {s} inside one step"
            self.r.setdefault(s, me)
        for s in writes:
            assert s not in self.w, f"WAW hazard on slot {s} inside one step"
            assert s not in self.r or self.r[s] == me, f"WAR hazard on slot {s} inside one step"
            self.w[s] = me


def _geometry(rec):
    f = int(rec.debug("stats")[14])            # frames per tile (32, or 16/8 for large factors)
    return f, 3 * f


def _parse_phases(rec):
    F, SLOT_WORDS = _geometry(rec)
    """Splits the stage stream into phases: list of levels, each a list of (target_slot, overwrite, coeff[], src_slot[])."""
    prog, stage_off = rec.debug("prog"), rec.debug("stage_off")
    phases, cur, level = [], None, []
    for s in range(len(stage_off) - 1):
        st = prog[stage_off[s]:stage_off[s + 1]]
        n_ops, nbytes = st[:8].view(np.uint32)
        assert nbytes == len(st) and nbytes <= 8192 and nbytes % 16 == 0
        at = 16
        for _ in range(n_ops):
            typ, flags = st[at:at + 4].view(np.uint16)
            a, b, c = st[at + 4:at + 16].view(np.uint32)
            if typ == OP_PHASE_BEGIN:
                assert cur is None
                cur, level = [], []
                at += 16
            elif typ == OP_PHASE_END:
                assert cur is not None and not level, "phase must end on a level barrier"
                phases.append(cur)
                cur = None
                at += 16
            else:
                assert typ == OP_ROWS and cur is not None
                table = st[b:b + 4 * a].view(np.uint32)
                for off in table:
                    t0, t1, t2, nf = st[off:off + 16].view(np.uint32)
                    n = int(nf & 0xFFFFFF)
                    assert off % 16 == 0
                    if nf & TASK_GROUP:                       # kind B: {src, c0, c1, c2} entries, up to 3 rows
                        assert n % 4 == 0
                        ent = st[off + 16: off + 16 + 16 * n]
                        src = ent.view(np.uint32)[0::4]
                        assert (src % (4 * SLOT_WORDS) == 0).all()
                        cf = ent.view(f32).reshape(n, 4)[:, 1:]
                        for r, tgt in enumerate((t0, t1, t2)):
                            if tgt == 0xFFFFFFFF:
                                assert not cf[:, r].any()
                                continue
                            assert tgt % (4 * SLOT_WORDS) == 0
                            level.append((int(tgt) // 4 // SLOT_WORDS, bool(nf & (TASK_OVERWRITE << r)), cf[:, r].copy(),
                                          (src // 4 // SLOT_WORDS).astype(np.int64)))
                        continue
                    tgt = t0
                    assert n % 8 == 0 and tgt % (4 * SLOT_WORDS) == 0
                    ent = st[off + 16: off + 16 + 8 * n]
                    src = ent.view(np.uint32)[1::2]
                    assert (src % (4 * SLOT_WORDS) == 0).all()
                    level.append((int(tgt) // 4 // SLOT_WORDS, bool(nf & TASK_OVERWRITE), ent.view(f32)[0::2].copy(),
                                  (src // 4 // SLOT_WORDS).astype(np.int64)))
                if flags & 2:
                    cur.append(level)
                    level = []
                at = int(c)
        assert at <= nbytes
    assert cur is None
    return phases


def solve(rec, rhs, cnst_pos=None):
    """K3 + K5: rhs [N, n_free, 3] f32 (permuted rows) -> verts [N, n_verts, 3] f32, following the kernel's
    real-time order of TMA loads, levels and stores (loads of phase q+2 are issued right after the stores of
    phase q), with NaN-poisoned shared memory and a slot-level race check inside every level."""
    stats = rec.debug("stats")
    F, SLOT_WORDS = _geometry(rec)
    n_slots = int(stats[0])
    io_desc = rec.debug("io_desc").reshape(-1, 4)
    io_phase = rec.debug("io_phase").reshape(-1, 4)
    phases = _parse_phases(rec)
    nf_, nb_ = int(stats[1]), len(io_phase) - int(stats[1])
    assert len(phases) == len(io_phase)
    perm, free_to_vi = rec.debug("perm"), rec.debug("free_to_vi")
    row_vert = free_to_vi[perm]
    xb = rec.debug("x_base").reshape(-1, 3)
    xb_hi = xb.astype(f32)
    xb_lo = (xb - xb_hi.astype(np.float64)).astype(f32)
    N = rhs.shape[0]
    out = np.full((N, rec.n_verts, 3), np.nan, dtype=f32)
    max_slot = 0
    for t0 in range(0, N, F):
        nv = min(F, N - t0)
        scratch = np.zeros((rec.n_free, 3, F), dtype=f32)         # tile-major rows; K2 zero-fills idle lanes
        scratch[:, :, :nv] = np.transpose(rhs[t0:t0 + nv], (1, 2, 0))
        state = np.full((n_slots, 3, F), np.nan, dtype=f32)
        busy = np.zeros(n_slots, dtype=bool)                       # slots some live phase may still touch

        def loads(q):
            for row, n, slot, _ in io_desc[io_phase[q][0]:io_phase[q][1]]:
                state[slot:slot + n] = scratch[row:row + n]

        def stores(q):
            for row, n, slot, _ in io_desc[io_phase[q][2]:io_phase[q][3]]:
                assert not np.isnan(state[slot:slot + n]).any()
                scratch[row:row + n] = state[slot:slot + n]

        def compute(q):
            nonlocal max_slot
            for level in phases[q]:
                hz = _Hazards()
                results = []
                for tgt, overwrite, coeff, src in level:
                    hz.access(set(int(x) for x, cf in zip(src, coeff) if cf != 0), {tgt})
                    max_slot = max(max_slot, tgt, int(src.max()) if len(src) else 0)
                    acc = np.zeros((3, F), dtype=f32)
                    for k in range(len(coeff)):
                        acc += coeff[k] * state[src[k]]
                    base = np.zeros((3, F), dtype=f32) if overwrite else state[tgt]
                    results.append((tgt, base - acc))
                for tgt, v in results:                              # all reads of a level precede its writes
                    state[tgt] = v

        q0 = 0
        for nq in (nf_, nb_):
            loads(q0)
            if nq > 1:
                loads(q0 + 1)
            for q in range(nq):
                compute(q0 + q)
                stores(q0 + q)
                if q + 2 < nq:
                    loads(q0 + q + 2)
            q0 += nq
        x = scratch[:, :, :nv]                                      # [row][c][frame]
        assert not np.isnan(x).any()
        val = xb_hi[:, :, None] + (xb_lo[:, :, None] + x)
        out[t0:t0 + nv, row_vert] = np.transpose(val, (2, 0, 1))
    if rec.n_cnsts:
        C = rec._verts[rec._cnsts] if cnst_pos is None else np.asarray(cnst_pos, dtype=f32).reshape(-1, 3)
        out[:, rec._cnsts] = C[None]
    return out, dict(max_slot_used=int(max_slot), n_slots=n_slots,
                     levels_per_tile=sum(len(p) for p in phases))
