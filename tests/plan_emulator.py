"""TEST-ONLY numpy interpreter of the plans that sdfa_create() uploads to the GPU.

It executes, in float32 and with the kernels' exact data flow, the assembly plan (kernel K2) and the
solve byte-program (kernel K3) fetched through ``sdfa_debug_get`` from a handle built with
``device=-1``.  This lets the CPU test tier validate the host-side scheduler (pieces, slots, levels,
op encoding) without a GPU.  It is NOT a fallback: nothing under ``sdfa-2019_b200/`` imports it.
"""
import numpy as np

F, CS = 32, 33                      # FRAMES_PER_TILE, COORD_STRIDE (csrc/plan.hpp)
SLOT_WORDS = 3 * CS
OP_ROWS, OP_LOAD, OP_STORE_Y, OP_STORE_X = 1, 2, 3, 4
LOAD_ADD = 0x80000000
TASK_FINAL, TASK_OVERWRITE = 1 << 24, 1 << 25
f32 = np.float32


def assemble(rec, dgrad, mode="dgrad"):
    """K2: dgrad [N, n_src*9] f32 -> rhs [N, n_free, 3] f32 (permuted row order)."""
    dg = np.ascontiguousarray(dgrad, dtype=f32).reshape(len(dgrad), -1, 9)
    blocks = rec.debug("asm_blocks").reshape(-1, 4)
    eq_id, eq_u = rec.debug("asm_eq_id"), rec.debug("asm_eq_u").reshape(-1, 6)
    row_perm, row_ptr, inc = rec.debug("asm_row_perm"), rec.debug("asm_row_ptr"), rec.debug("asm_inc")
    eq_src = rec.debug("eq_src")
    N = dg.shape[0]
    rhs = np.zeros((N, rec.n_free, 3), dtype=f32)
    for eb, ee, rb, re in blocks:
        ids = eq_id[eb:ee]
        src = eq_src[ids]
        u0, u1 = eq_u[eb:ee, :3], eq_u[eb:ee, 3:]
        g = np.zeros((N, ee - eb, 3, 3), dtype=f32)          # [frame][eq][corner][xyz]
        has = src >= 0
        if has.any():
            d = dg[:, src[has]]                               # [N, m, 9]
            if mode == "dgrad":
                th2 = d[..., 6] ** 2 + d[..., 7] ** 2 + d[..., 8] ** 2
                th = np.sqrt(th2)
                ok = th >= f32(1e-6)
                ths = np.where(ok, th, f32(1))
                a = np.where(ok, np.sin(ths) / ths, f32(0)).astype(f32)
                b = np.where(ok, f32(2) * np.sin(f32(0.5) * ths) ** 2 / np.where(ok, th2, f32(1)), f32(0)).astype(f32)

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
        g[:, :, 0] = -(g[:, :, 1] + g[:, :, 2])
        gf = g.reshape(N, -1, 3)
        for r in range(rb, re):
            idx = inc[row_ptr[r]:row_ptr[r + 1]].astype(np.int64)
            acc = np.zeros((N, 3), dtype=f32)
            for i in idx:                                     # same sequential order as the kernel
                acc = acc + gf[:, i]
            rhs[:, row_perm[r]] = acc
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


def solve(rec, rhs, cnst_pos=None):
    """K3 + K4: rhs [N, n_free, 3] f32 -> verts [N, n_verts, 3] f32.  Also returns bookkeeping stats."""
    prog, stage_off = rec.debug("prog"), rec.debug("stage_off")
    stats = rec.debug("stats")
    n_slots = int(stats[0])
    perm, free_to_vi = rec.debug("perm"), rec.debug("free_to_vi")
    row_vert = free_to_vi[perm]
    xb = rec.debug("x_base").reshape(-1, 3)
    xb_hi = xb.astype(f32)
    xb_lo = (xb - xb_hi.astype(np.float64)).astype(f32)
    N = rhs.shape[0]
    out = np.full((N, rec.n_verts, 3), np.nan, dtype=f32)
    scratch = np.array(rhs, dtype=f32, copy=True)
    max_slot_used = 0
    n_sync = 0
    for t0 in range(0, N, F):
        nv = min(F, N - t0)
        state = np.zeros(n_slots * SLOT_WORDS, dtype=f32)
        hz = _Hazards()
        lanes = np.arange(nv)
        for s in range(len(stage_off) - 1):
            st = prog[stage_off[s]:stage_off[s + 1]]
            n_ops, nbytes = st[:8].view(np.uint32)
            assert nbytes == len(st) and nbytes <= 8192 and nbytes % 16 == 0
            at = 16
            for _ in range(n_ops):
                hdr = st[at:at + 16]
                typ, flags = hdr[:4].view(np.uint16)
                a, b, c = hdr[4:16].view(np.uint32)
                n_sync += bool(flags & 1) + bool(flags & 2)
                if flags & 1:
                    hz.sync()
                if typ == OP_ROWS:
                    table = st[b:b + 4 * a].view(np.uint32)
                    written = set()
                    for off in table:
                        tgt, nf, dinv_bits, _ = st[off:off + 16].view(np.uint32)
                        n = int(nf & 0xFFFFFF)
                        assert n % 2 == 0 and off % 16 == 0
                        ent = st[off + 16: off + 16 + 8 * n]
                        coeff = ent.view(f32)[0::2]
                        src = ent.view(np.uint32)[1::2]
                        assert tgt % 4 == 0 and tgt // 4 % SLOT_WORDS == 0
                        assert tgt not in written, "two tasks of one step write the same row"
                        written.add(int(tgt))
                        reads = set(int(x) // 4 // SLOT_WORDS for x, cf in zip(src, coeff) if cf != 0)
                        hz.access(reads, {int(tgt) // 4 // SLOT_WORDS})
                        max_slot_used = max(max_slot_used, tgt // 4 // SLOT_WORDS, *(src // 4 // SLOT_WORDS)) if n else max_slot_used
                        dinv = np.array([dinv_bits], dtype=np.uint32).view(f32)[0]
                        acc = np.zeros((3, nv), dtype=f32)
                        for k in range(n):
                            base = src[k] // 4
                            for cdim in range(3):
                                acc[cdim] += coeff[k] * state[base + cdim * CS + lanes]
                        for cdim in range(3):
                            w = tgt // 4 + cdim * CS + lanes
                            v = np.zeros(nv, dtype=f32) if (nf & TASK_OVERWRITE) else state[w]
                            state[w] = (v - acc[cdim]) * dinv
                    at = int(c)
                else:
                    table = st[c:c + 4 * b].view(np.uint32)
                    rows = np.arange(a, a + b)
                    for r, w in zip(rows, table):
                        base = int(w & 0xFFFFFF)
                        max_slot_used = max(max_slot_used, base // SLOT_WORDS)
                        for cdim in range(3):
                            idx = base + cdim * CS + lanes
                            if typ == OP_LOAD:
                                v = scratch[t0:t0 + nv, r, cdim]
                                state[idx] = (state[idx] + v) if (w & LOAD_ADD) else v
                            elif typ == OP_STORE_Y:
                                scratch[t0:t0 + nv, r, cdim] = state[idx]
                            elif typ == OP_STORE_X:
                                out[t0:t0 + nv, row_vert[r], cdim] = xb_hi[r, cdim] + (xb_lo[r, cdim] + state[idx])
                            else:
                                raise AssertionError(f"bad op {typ}")
                    slots = set(int(w & 0xFFFFFF) // SLOT_WORDS for w in table)
                    # every warp touches every row of a span op (frames are dealt to warps), so the op
                    # conflicts with any other unsynchronised access to those slots
                    if typ == OP_LOAD:
                        hz.access(set(), slots)
                    else:
                        hz.access(slots, set())
                    at = int(c) + (int(b) * 4 + 15) // 16 * 16
                if flags & 2:
                    hz.sync()
            assert at <= nbytes
    if rec.n_cnsts:
        C = rec._verts[rec._cnsts] if cnst_pos is None else np.asarray(cnst_pos, dtype=f32).reshape(-1, 3)
        out[:, rec._cnsts] = C[None]
    return out, dict(max_slot_used=int(max_slot_used), n_slots=n_slots, syncs_per_tile=n_sync // max(1, (N + F - 1) // F))
