// schedule.cpp -- turns the Cholesky factor into the byte program of the solve kernel (K3) and builds
// the row-block plan of the assembly kernel (K2).  Pure host code; what it emits is documented in plan.hpp.
//
// Solve schedule.  The permuted system is walked in "pieces": contiguous ranges [a,b) of the postordered
// elimination tree.  Because every column's pattern is a subset of its ancestors, the only rows outside
// a piece that a piece touches lie on the root path above it, so a CTA never needs more than
// piece + path rows resident in shared memory while it streams the rest through global scratch:
//   forward  (L y = b):   load rhs rows of the piece, run its rows level by level (row i subtracts
//                         L_ij y_j over the columns j inside the piece), push the piece's contribution into
//                         the not-yet-final rows above it (partial-row tasks, one per touched ancestor,
//                         so no atomics), spill y of the piece;
//   backward (L^T x = y): pieces in reverse; x_j needs x_i only for rows i in column j's pattern, all
//                         either inside the piece or still resident from the pieces above.
// This replaces solver_.solve() of the reference (deform_triangle_impl.hpp:286; Eigen SparseLU.h:217-241).
#include "plan.hpp"

#include <algorithm>
#include <cassert>
#include <cstring>
#include <map>
#include <numeric>
#include <set>
#include <stdexcept>
#include <cstddef>
#include <cstdlib>

namespace sdfa {

namespace {

struct Task {
    int target_slot;
    uint32_t flags;
    std::vector<TaskEntry> entries;     // src_byte_off holds a (symbolic) slot id until emission
    int target_row = -1;                // permuted row the task writes (grouping key)
};

// A serialised task (header + entries) ready to be placed into a stage.
struct Blob {
    std::vector<uint8_t> bytes;
    size_t work;                        // multiply-adds per lane and coordinate, for load balancing
};

class Emitter {
public:
    explicit Emitter(SolveProgram &prog) : prog_(prog) { open_stage(); }

    // one level: independent tasks, dealt round-robin to the consumer warps by the kernel
    void rows(std::vector<Blob> &tasks, bool sync_after) {
        std::stable_sort(tasks.begin(), tasks.end(), [](const Blob &a, const Blob &b) { return a.work > b.work; });
        size_t i = 0;
        while (i < tasks.size()) {
            size_t room = STAGE_BYTES - cur_.size();
            size_t n = 0, bytes = sizeof(OpHeader);
            while (i + n < tasks.size()) {
                size_t table = ((n + 1) * 4 + 15) / 16 * 16;
                size_t body = tasks[i + n].bytes.size();
                size_t prev_table = (n * 4 + 15) / 16 * 16;
                if (bytes - prev_table + table + body > room) break;
                bytes = bytes - prev_table + table + body;
                ++n;
            }
            if (n == 0) {
                if (n_ops_ == 0) throw std::runtime_error("solve schedule: a row task is larger than a stage");
                close_stage();
                open_stage();
                continue;
            }
            bool last = (i + n == tasks.size());
            OpHeader h{OP_ROWS, (uint16_t)((last && sync_after) ? OPF_SYNC_AFTER : 0), (uint32_t)n, 0, 0};
            size_t op_at = cur_.size();
            size_t table_at = op_at + sizeof(OpHeader);
            size_t table_bytes = (n * 4 + 15) / 16 * 16;
            h.b = (uint32_t)table_at;
            cur_.resize(table_at + table_bytes, 0);
            std::memcpy(&cur_[op_at], &h, sizeof(h));
            for (size_t t = 0; t < n; ++t) {
                uint32_t off = (uint32_t)cur_.size();
                std::memcpy(&cur_[table_at + t * 4], &off, 4);
                cur_.insert(cur_.end(), tasks[i + t].bytes.begin(), tasks[i + t].bytes.end());
            }
            uint32_t next = (uint32_t)cur_.size();
            std::memcpy(&cur_[op_at + offsetof(OpHeader, c)], &next, 4);
            ++n_ops_;
            i += n;
        }
    }

    void marker(OpType type) {
        if (cur_.size() + sizeof(OpHeader) > STAGE_BYTES) { close_stage(); open_stage(); }
        OpHeader h{(uint16_t)type, 0, 0, 0, 0};
        size_t op_at = cur_.size();
        cur_.resize(op_at + sizeof(OpHeader), 0);
        std::memcpy(&cur_[op_at], &h, sizeof(h));
        ++n_ops_;
    }

    void finish() {
        close_stage();
        prog_.stage_off.push_back((uint32_t)prog_.bytes.size());
    }

private:
    void open_stage() {
        cur_.assign(sizeof(StageHeader), 0);
        n_ops_ = 0;
    }
    void close_stage() {
        if (n_ops_ == 0) return;
        StageHeader sh{(uint32_t)n_ops_, (uint32_t)cur_.size(), 0, 0};
        std::memcpy(&cur_[0], &sh, sizeof(sh));
        prog_.stage_off.push_back((uint32_t)prog_.bytes.size());
        prog_.bytes.insert(prog_.bytes.end(), cur_.begin(), cur_.end());
        n_ops_ = 0;
    }
    SolveProgram &prog_;
    std::vector<uint8_t> cur_;
    int n_ops_ = 0;
};

// Slot allocator with the reuse delay the kernel's prefetching needs: the loads of phase q+2 are issued
// while phase q+1 computes, so a slot released at the end of phase q may be handed out again no earlier
// than for phase q+2.
class SlotPool {
public:
    // contiguous range if one exists below the high-water mark, else grows the pool
    int alloc_range(int n) {
        int run = 0;
        for (int s = 0; s < (int)used_.size(); ++s) {
            run = used_[s] ? 0 : run + 1;
            if (run == n) { int b = s - n + 1; mark(b, n); return b; }
        }
        int b = (int)used_.size() - run;               // extend a free tail
        used_.resize(b + n, 0);
        mark(b, n);
        return b;
    }
    int alloc() { return alloc_range(1); }
    void release_at_end_of(int phase, int slot) { pending_.push_back({phase, slot}); }
    // make everything released in phases <= phase available
    void reclaim(int phase) {
        size_t k = 0;
        for (auto &pr : pending_) {
            if (pr.first <= phase) used_[pr.second] = 0;
            else pending_[k++] = pr;
        }
        pending_.resize(k);
    }
    int peak() const { return (int)used_.size(); }
private:
    void mark(int b, int n) { for (int s = b; s < b + n; ++s) used_[s] = 1; }
    std::vector<char> used_;
    std::vector<std::pair<int, int>> pending_;
};

// merge (row, slot) pairs into runs of consecutive rows landing in consecutive slots
void make_descs(std::vector<std::pair<int, int>> &rows_slots, std::vector<IoDesc> &out) {
    std::sort(rows_slots.begin(), rows_slots.end());
    for (size_t i = 0; i < rows_slots.size();) {
        size_t j = i + 1;
        while (j < rows_slots.size() && rows_slots[j].first == rows_slots[j - 1].first + 1 &&
               rows_slots[j].second == rows_slots[j - 1].second + 1)
            ++j;
        out.push_back({(uint32_t)rows_slots[i].first, (uint32_t)(j - i), (uint32_t)rows_slots[i].second, 0});
        i = j;
    }
}

// ---- serialisation of tasks (formats in plan.hpp) --------------------------------------------------
// kind A: one target row, entries {coeff, src} packed two per 16 bytes.
Blob blob_single(const Task &t, int SLOT_BYTES) {
    Blob b;
    // padded to a multiple of 8 entries (4 packed pairs) so the kernel's loop has no remainder iterations
    const size_t n = t.entries.size(), ne = (n + 7) / 8 * 8;
    b.bytes.assign(sizeof(TaskHeader) + ne * sizeof(TaskEntry), 0);
    TaskHeader th{(uint32_t)t.target_slot * SLOT_BYTES, 0, 0, (uint32_t)ne | t.flags};
    std::memcpy(&b.bytes[0], &th, sizeof(th));
    for (size_t k = 0; k < n; ++k) {
        TaskEntry e{t.entries[k].coeff, t.entries[k].src_byte_off * SLOT_BYTES};
        std::memcpy(&b.bytes[sizeof(TaskHeader) + k * sizeof(TaskEntry)], &e, sizeof(e));
    }
    for (size_t k = n; k < ne; ++k) {     // padding: coefficient 0 on the task's own first source (a live, finite slot)
        TaskEntry e{0.f, t.entries[0].src_byte_off * SLOT_BYTES};
        std::memcpy(&b.bytes[sizeof(TaskHeader) + k * sizeof(TaskEntry)], &e, sizeof(e));
    }
    b.work = ne;
    return b;
}
// kind B: up to three target rows that read (nearly) the same sources: one entry {src, c0, c1, c2} per
// source, so each source value is fetched from shared memory once for all rows of the group.
Blob blob_group(const std::vector<const Task *> &rows, int SLOT_BYTES) {
    std::map<uint32_t, GroupEntry> uni;
    for (size_t r = 0; r < rows.size(); ++r)
        for (const TaskEntry &e : rows[r]->entries) {
            GroupEntry &g = uni[e.src_byte_off];
            g.src_byte_off = e.src_byte_off * SLOT_BYTES;
            g.c[r] += e.coeff;
        }
    Blob b;
    const size_t ne = (uni.size() + TASK_BATCH_B - 1) / TASK_BATCH_B * TASK_BATCH_B;   // whole batches (zero coefficients)
    b.bytes.assign(sizeof(TaskHeader) + ne * sizeof(GroupEntry), 0);
    TaskHeader th{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)ne | TASK_GROUP};
    uint32_t *tg = &th.target_byte_off;
    for (size_t r = 0; r < rows.size(); ++r) {
        tg[r] = (uint32_t)rows[r]->target_slot * SLOT_BYTES;
        if (rows[r]->flags & TASK_OVERWRITE) th.n_entries_flags |= (TASK_OVERWRITE << r);
    }
    std::memcpy(&b.bytes[0], &th, sizeof(th));
    size_t k = 0;
    for (auto &kv : uni) std::memcpy(&b.bytes[sizeof(TaskHeader) + (k++) * sizeof(GroupEntry)], &kv.second, sizeof(GroupEntry));
    for (; k < ne; ++k) {
        GroupEntry pad{uni.begin()->second.src_byte_off, {0.f, 0.f, 0.f}};
        std::memcpy(&b.bytes[sizeof(TaskHeader) + k * sizeof(GroupEntry)], &pad, sizeof(pad));
    }
    b.work = ne * 3;
    return b;
}

// Groups the tasks of one level: neighbouring target rows (same block of the factor) whose source lists
// nearly coincide share one kind-B task.  `pad_limit` bounds the zero padding the union may introduce.
std::vector<Blob> group_level(std::vector<Task> &lv, const std::vector<int> &block_of_row, double pad_limit,
                              long long &useful, long long &padded, int SLOT_BYTES) {
    std::vector<Blob> out;
    std::stable_sort(lv.begin(), lv.end(), [&](const Task &a, const Task &b) { return a.target_row < b.target_row; });
    size_t i = 0;
    while (i < lv.size()) {
        std::vector<const Task *> grp{&lv[i]};
        std::set<uint32_t> uni;
        for (auto &e : lv[i].entries) uni.insert(e.src_byte_off);
        size_t sum = lv[i].entries.size(), j = i + 1;
        while (j < lv.size() && grp.size() < 3 && block_of_row[lv[j].target_row] == block_of_row[lv[i].target_row]) {
            std::set<uint32_t> u2 = uni;
            for (auto &e : lv[j].entries) u2.insert(e.src_byte_off);
            const size_t sum2 = sum + lv[j].entries.size();
            // cost of the group = |union| * rows; accept while the padding stays small
            if ((double)(u2.size() * (grp.size() + 1)) > pad_limit * (double)sum2) break;
            uni.swap(u2);
            sum = sum2;
            grp.push_back(&lv[j]);
            ++j;
        }
        useful += (long long)sum;
        if (grp.size() == 1) { out.push_back(blob_single(lv[i], SLOT_BYTES)); padded += (long long)sum; }
        else { out.push_back(blob_group(grp, SLOT_BYTES)); padded += (long long)(uni.size() * 3); }
        i = j;
    }
    return out;
}

struct Supernode {
    int c0, c1;                    // columns [c0, c1)
    std::vector<int> below;        // rows below the diagonal block (pattern of the last column)
    std::vector<double> linv;      // w x w row-major: inverse of the dense diagonal block (lower triangular)
    std::vector<double> wmat;      // h x w row-major: L[below, J] * inv(L[J, J])
    int parent = -1;
};

}  // namespace

// Blocks ("supernodes") of consecutive columns of the postordered factor, of two kinds:
//   * chains: maximal runs j, j+1, ... where j+1 is j's parent and only child and pattern(j+1) =
//     pattern(j) \ {j+1} (fundamental supernodes: dense diagonal block), cut at `cap` columns;
//   * whole small subtrees (<= subtree_cap columns): contiguous in postorder as well, sparse inside.
// Inverting the small diagonal block of a block J on the host (fp64) removes every dependency INSIDE it:
//      y_J = inv(L_JJ) t_J,     t_I -= (L_IJ inv(L_JJ)) t_J          (forward,  I = rows below J)
//      x_J = inv(L_JJ)^T y_J - (L_IJ inv(L_JJ))^T x_I                (backward)
// so a sweep needs one barrier per level of the BLOCK tree instead of one per column.  inv(L_JJ) of a
// subtree block is only as dense as the ancestor relation inside the subtree; exact zeros are dropped.
static std::vector<Supernode> find_supernodes(const HostPlan &p, int cap, int subtree_cap) {
    const int n = p.n_free;
    std::vector<int> nchild(n, 0), size(n, 1);
    for (int j = 0; j < n; ++j)
        if (p.parent[j] >= 0) { nchild[p.parent[j]]++; size[p.parent[j]] += size[j]; }
    auto cnt = [&](int j) { return p.l_colptr[j + 1] - p.l_colptr[j]; };
    // A pivot that is ~reg next to the matrix scale marks a direction the constraints do not fix (no
    // constraints: one translation per connected component).  The fp64 reference resolves it
    // through the 1e-10 regulariser into an arbitrary, noise-driven translation; float32 sweeps cannot
    // (1/pivot ~ 1e5 amplifies rounding), so the displacement of that row is pinned to zero instead
    // (its inverse pivot is taken as 0) -- the result differs from the reference by exactly that gauge.
    double max_diag = 0.0;
    for (int c = 0; c < n; ++c)
        for (int q = p.m_colptr[c]; q < p.m_colptr[c + 1]; ++q)
            if (p.m_rowidx[q] == c) max_diag = std::max(max_diag, p.m_val[q]);
    const double tiny_pivot2 = 1e-9 * max_diag;
    // block boundaries
    std::vector<std::pair<int, int>> ranges;
    for (int j = 0; j < n;) {
        // A leaf j may start a small subtree: climb while the parent's subtree also starts at j and fits.
        if (size[j] == 1) {
            int best = j, r = j;
            while (true) {
                int pr = p.parent[r];
                if (pr < 0 || pr - size[pr] + 1 != j || size[pr] > subtree_cap) break;
                best = r = pr;
            }
            if (best > j) { ranges.push_back({j, best + 1}); j = best + 1; continue; }
        }
        // chain supernode starting at j
        int e = j + 1;
        while (e < n && p.parent[e - 1] == e && nchild[e] == 1 && cnt(e) == cnt(e - 1) - 1 && (e - j) < cap) ++e;
        ranges.push_back({j, e});
        j = e;
    }
    std::vector<Supernode> sn;
    for (auto [start, j] : ranges) {
        Supernode s;
        s.c0 = start; s.c1 = j;
        const int w = j - start;
        for (int col = start; col < j; ++col)
            for (int q = p.l_colptr[col] + 1; q < p.l_colptr[col + 1]; ++q)
                if (p.l_rowidx[q] >= j) s.below.push_back(p.l_rowidx[q]);
        std::sort(s.below.begin(), s.below.end());
        s.below.erase(std::unique(s.below.begin(), s.below.end()), s.below.end());
        const int h = (int)s.below.size();
        // diagonal block T (lower) and block below B, as dense arrays
        std::vector<double> T((size_t)w * w, 0.0), B((size_t)h * w, 0.0);
        for (int b = 0; b < w; ++b) {
            int col = start + b;
            for (int q = p.l_colptr[col]; q < p.l_colptr[col + 1]; ++q) {
                int r = p.l_rowidx[q];
                if (r < j) T[(size_t)(r - start) * w + b] = p.l_val[q];
                else {
                    int k = (int)(std::lower_bound(s.below.begin(), s.below.end(), r) - s.below.begin());
                    B[(size_t)k * w + b] = p.l_val[q];
                }
            }
        }
        // inverse of T by forward substitution on the identity
        s.linv.assign((size_t)w * w, 0.0);
        for (int c = 0; c < w; ++c)
            for (int r = c; r < w; ++r) {
                double v = (r == c) ? 1.0 : 0.0;
                for (int k = c; k < r; ++k) v -= T[(size_t)r * w + k] * s.linv[(size_t)k * w + c];
                const double piv = T[(size_t)r * w + r];
                s.linv[(size_t)r * w + c] = (piv * piv < tiny_pivot2) ? 0.0 : v / piv;
            }
        s.wmat.assign((size_t)h * w, 0.0);
        for (int k = 0; k < h; ++k)
            for (int c = 0; c < w; ++c) {
                double v = 0.0;
                for (int b = c; b < w; ++b) v += B[(size_t)k * w + b] * s.linv[(size_t)b * w + c];
                s.wmat[(size_t)k * w + c] = v;
            }
        sn.push_back(std::move(s));
    }
    std::vector<int> sn_of(n);
    for (int i = 0; i < (int)sn.size(); ++i) for (int c = sn[i].c0; c < sn[i].c1; ++c) sn_of[c] = i;
    for (auto &s : sn) { int pr = p.parent[s.c1 - 1]; s.parent = pr >= 0 ? sn_of[pr] : -1; }
    return sn;
}

namespace {
// Slots come from two pools so that long-lived rows cannot fragment the space the short-lived piece
// blocks rotate through: "ext" slots (rows that outlive the phase that first touches them -- partial
// sums of ancestors in the forward sweep, solved ancestors in the backward sweep) and "piece" slots
// (released at the end of their phase).  Ids are symbolic until both sweeps are simulated and the size
// of the ext pool is known: final slot = ext ? id : n_ext + id.
constexpr int EXT_FLAG = 1 << 30;

struct PhaseData {
    std::vector<std::vector<Task>> levels;
    std::vector<std::pair<int, int>> loads, stores;     // (row, symbolic slot)
};
}  // namespace

void build_solve_program(HostPlan &p, int piece_cap, int supernode_cap, int subtree_cap, int frames_per_tile) {
    const int n = p.n_free;
    SolveProgram &prog = p.prog;
    prog = SolveProgram();
    prog.frames_per_tile = frames_per_tile;
    const int SLOT_BYTES = slot_bytes(frames_per_tile);
    std::vector<Supernode> sn = find_supernodes(p, supernode_cap, subtree_cap);
    const int ns = (int)sn.size();
    prog.n_supernodes = ns;
    // pieces: consecutive supernodes, about piece_cap rows each
    std::vector<std::pair<int, int>> pieces;       // supernode ranges [s0, s1)
    for (int s0 = 0; s0 < ns;) {
        int s1 = s0, rows = 0;
        while (s1 < ns && (rows == 0 || rows + (sn[s1].c1 - sn[s1].c0) <= piece_cap)) { rows += sn[s1].c1 - sn[s1].c0; ++s1; }
        pieces.push_back({s0, s1});
        s0 = s1;
    }
    const int np = (int)pieces.size();
    std::vector<int> piece_of_row(n);
    for (int pi = 0; pi < np; ++pi)
        for (int s = pieces[pi].first; s < pieces[pi].second; ++s)
            for (int c = sn[s].c0; c < sn[s].c1; ++c) piece_of_row[c] = pi;
    std::vector<int> lvl(ns, 0);
    std::vector<PhaseData> phases;
    int n_ext = 0, n_piece = 0;

    // ---------------------------------------------------------------- forward sweep
    {
        SlotPool pool_p, pool_e;
        std::vector<int> tslot(n, -1), yslot(n, -1);
        auto prepare = [&](int q) {
            const int s0 = pieces[q].first, s1 = pieces[q].second;
            const int a = sn[s0].c0, b = sn[s1 - 1].c1;
            PhaseData ph;
            // t-slots: rows of the piece that no earlier piece has touched yet, as one contiguous block
            std::vector<int> fresh;
            for (int i = a; i < b; ++i) if (tslot[i] < 0) fresh.push_back(i);
            if (!fresh.empty()) {
                int base = pool_p.alloc_range((int)fresh.size());
                for (size_t k = 0; k < fresh.size(); ++k) { tslot[fresh[k]] = base + (int)k; ph.loads.push_back({fresh[k], base + (int)k}); }
            }
            int ybase = pool_p.alloc_range(b - a);
            for (int i = a; i < b; ++i) { yslot[i] = ybase + (i - a); ph.stores.push_back({i, yslot[i]}); }
            // rows above the piece that receive a contribution for the first time: bring in their rhs
            for (int s = s0; s < s1; ++s)
                for (int i : sn[s].below)
                    if (i >= b && tslot[i] < 0) { tslot[i] = EXT_FLAG | pool_e.alloc(); ph.loads.push_back({i, tslot[i]}); }
            // tasks, by level of the supernodal tree inside the piece.  The piece's pushes into a row above
            // its supernodes are merged into ONE task per target row: it reads the t-values of every
            // contributing supernode (they stay intact, y goes to separate slots) and runs in the level
            // of the last contributor -- always before the level of the target's own supernode.
            std::map<int, Task> ext;
            std::map<int, int> ext_lvl;
            for (int s = s0; s < s1; ++s) lvl[s] = 0;
            for (int s = s0; s < s1; ++s) {
                const Supernode &S = sn[s];
                const int l = lvl[s], w = S.c1 - S.c0;
                if (S.parent >= 0 && S.parent < s1) lvl[S.parent] = std::max(lvl[S.parent], l + 1);
                if ((int)ph.levels.size() <= l) ph.levels.resize(l + 1);
                for (int r = 0; r < w; ++r) {              // y_i = sum_b linv[r][b] t_b  (target overwritten)
                    Task t{yslot[S.c0 + r], TASK_OVERWRITE, {}, S.c0 + r};
                    for (int c = 0; c <= r; ++c)
                        if (S.linv[(size_t)r * w + c] != 0.0 || c == r)
                            t.entries.push_back({(float)(-S.linv[(size_t)r * w + c]), (uint32_t)tslot[S.c0 + c]});
                    ph.levels[l].push_back(std::move(t));
                }
                for (size_t k = 0; k < S.below.size(); ++k) {   // t_i -= sum_b W[k][b] t_b
                    int i = S.below[k];
                    auto it = ext.find(i);
                    if (it == ext.end()) { it = ext.emplace(i, Task{tslot[i], 0, {}, i}).first; ext_lvl[i] = 0; }
                    for (int c = 0; c < w; ++c)
                        if (S.wmat[k * w + c] != 0.0)
                            it->second.entries.push_back({(float)S.wmat[k * w + c], (uint32_t)tslot[S.c0 + c]});
                    ext_lvl[i] = std::max(ext_lvl[i], l);
                }
            }
            for (auto &kv : ext) ph.levels[ext_lvl[kv.first]].push_back(std::move(kv.second));
            prog.n_steps_fwd += (int)ph.levels.size();
            phases.push_back(std::move(ph));
        };
        auto retire = [&](int q) {
            const int a = sn[pieces[q].first].c0, b = sn[pieces[q].second - 1].c1;
            for (int i = a; i < b; ++i) {
                if (tslot[i] & EXT_FLAG) pool_e.release_at_end_of(q, tslot[i] & ~EXT_FLAG);
                else pool_p.release_at_end_of(q, tslot[i]);
                pool_p.release_at_end_of(q, yslot[i]);
            }
        };
        prepare(0);
        if (np > 1) prepare(1);
        for (int q = 0; q < np; ++q) {
            retire(q);
            if (q + 2 < np) { pool_p.reclaim(q); pool_e.reclaim(q); prepare(q + 2); }
        }
        prog.n_phases_fwd = np;
        n_ext = std::max(n_ext, pool_e.peak());
        n_piece = std::max(n_piece, pool_p.peak());
    }
    // ---------------------------------------------------------------- backward sweep
    {
        SlotPool pool_p, pool_e;
        std::vector<int> yslot(n, -1), xslot(n, -1);
        // x of row i stays resident until the piece holding the first column that has i below it is done
        std::vector<int> last_use(n);
        for (int i = 0; i < n; ++i) last_use[i] = piece_of_row[i];
        for (int s = 0; s < ns; ++s)
            for (int i : sn[s].below) last_use[i] = std::min(last_use[i], piece_of_row[sn[s].c0]);
        std::vector<std::vector<int>> release_after(np);
        for (int i = 0; i < n; ++i) release_after[last_use[i]].push_back(i);
        auto prepare = [&](int q) {                 // phase q of the sweep handles piece np-1-q
            const int pi = np - 1 - q;
            const int s0 = pieces[pi].first, s1 = pieces[pi].second;
            const int a = sn[s0].c0, b = sn[s1 - 1].c1;
            PhaseData ph;
            int ybase = pool_p.alloc_range(b - a);
            std::vector<int> shortlived;
            for (int i = a; i < b; ++i) {
                yslot[i] = ybase + (i - a);
                ph.loads.push_back({i, yslot[i]});
                if (last_use[i] == pi) shortlived.push_back(i);
                else xslot[i] = EXT_FLAG | pool_e.alloc();
            }
            if (!shortlived.empty()) {
                int xbase = pool_p.alloc_range((int)shortlived.size());
                for (size_t k = 0; k < shortlived.size(); ++k) xslot[shortlived[k]] = xbase + (int)k;
            }
            for (int i = a; i < b; ++i) ph.stores.push_back({i, xslot[i]});
            for (int s = s1 - 1; s >= s0; --s) {
                const Supernode &S = sn[s];
                const int w = S.c1 - S.c0;
                const int l = (S.parent >= 0 && S.parent < s1) ? lvl[S.parent] + 1 : 0;
                lvl[s] = l;
                if ((int)ph.levels.size() <= l) ph.levels.resize(l + 1);
                for (int c = 0; c < w; ++c) {              // x_j = sum_r linv[r][c] y_r - sum_k W[k][c] x_below[k]
                    Task t{xslot[S.c0 + c], TASK_OVERWRITE, {}, S.c0 + c};
                    for (int r = c; r < w; ++r)
                        if (S.linv[(size_t)r * w + c] != 0.0 || r == c)
                            t.entries.push_back({(float)(-S.linv[(size_t)r * w + c]), (uint32_t)yslot[S.c0 + r]});
                    for (size_t k = 0; k < S.below.size(); ++k) {
                        assert(xslot[S.below[k]] >= 0);
                        if (S.wmat[k * w + c] != 0.0)
                            t.entries.push_back({(float)S.wmat[k * w + c], (uint32_t)xslot[S.below[k]]});
                    }
                    ph.levels[l].push_back(std::move(t));
                }
            }
            prog.n_steps_bwd += (int)ph.levels.size();
            phases.push_back(std::move(ph));
        };
        auto retire = [&](int q) {
            const int pi = np - 1 - q;
            const int a = sn[pieces[pi].first].c0, b = sn[pieces[pi].second - 1].c1;
            for (int i = a; i < b; ++i) pool_p.release_at_end_of(q, yslot[i]);
            for (int i : release_after[pi]) {
                if (xslot[i] & EXT_FLAG) pool_e.release_at_end_of(q, xslot[i] & ~EXT_FLAG);
                else pool_p.release_at_end_of(q, xslot[i]);
            }
        };
        prepare(0);
        if (np > 1) prepare(1);
        for (int q = 0; q < np; ++q) {
            retire(q);
            if (q + 2 < np) { pool_p.reclaim(q); pool_e.reclaim(q); prepare(q + 2); }
        }
        prog.n_phases_bwd = np;
        n_ext = std::max(n_ext, pool_e.peak());
        n_piece = std::max(n_piece, pool_p.peak());
    }
    // ---------------------------------------------------------------- resolve slots and emit
    prog.n_slots = n_ext + n_piece;
    auto resolve = [&](int sym) { return (sym & EXT_FLAG) ? (sym & ~EXT_FLAG) : n_ext + sym; };
    Emitter em(prog);
    std::vector<int> block_of_row(n);
    for (int b = 0; b < ns; ++b) for (int c = sn[b].c0; c < sn[b].c1; ++c) block_of_row[c] = b;
    const char *pl = std::getenv("SDFA_GROUP_PAD");
    const double pad_limit = pl ? std::atof(pl) : 1.5;
    long long useful = 0, padded = 0;
    for (auto &ph : phases) {
        for (auto &rs : ph.loads) rs.second = resolve(rs.second);
        for (auto &rs : ph.stores) rs.second = resolve(rs.second);
        IoPhase io;
        io.load_begin = (uint32_t)prog.io_desc.size();
        make_descs(ph.loads, prog.io_desc);
        io.load_end = io.store_begin = (uint32_t)prog.io_desc.size();
        make_descs(ph.stores, prog.io_desc);
        io.store_end = (uint32_t)prog.io_desc.size();
        prog.io_phase.push_back(io);
        em.marker(OP_PHASE_BEGIN);
        for (auto &lv : ph.levels) {
            for (auto &t : lv) {
                t.target_slot = resolve(t.target_slot);
                for (auto &e : t.entries) e.src_byte_off = (uint32_t)resolve((int)e.src_byte_off);
            }
            std::vector<Blob> blobs = group_level(lv, block_of_row, pad_limit, useful, padded, SLOT_BYTES);
            em.rows(blobs, true);
        }
        em.marker(OP_PHASE_END);
    }
    em.finish();
    prog.n_entries = useful;
    prog.n_entries_padded = padded;
}

// ------------------------------------------------------------------------------------------
void build_assembly_plan(HostPlan &p, int rows_per_block, int max_eq_per_block) {
    AssemblyPlan &ap = p.asmplan;
    ap = AssemblyPlan();
    const int n = p.n_free;
    // incidences per free column: (equation block, corner)
    std::vector<std::vector<std::pair<int, int>>> inc(n);
    std::vector<std::vector<int>> rows_of_eq(p.n_eq);
    for (int k : p.active_eq) {
        const uint32_t *t = &p.tris[(size_t)p.eq_tri[k] * 3];
        for (int c = 0; c < 3; ++c) {
            int f = p.vi_to_free[t[c]];
            if (f >= 0) { inc[f].push_back({k, c}); rows_of_eq[k].push_back(f); }
        }
    }
    std::vector<int> key(n, 0x7fffffff);
    for (int f = 0; f < n; ++f)
        for (auto &kc : inc[f]) key[f] = std::min(key[f], kc.first);
    // Row blocks by greedy growth over the mesh: start at the unassigned row with the smallest equation
    // index and keep adding the neighbouring row that brings the fewest new equations, so that few
    // equations are evaluated by more than one block (FLAME: 1.17x instead of 1.67x for index-sorted rows).
    ap.row_ptr.push_back(0);
    std::vector<char> taken(n, 0), in_eqs(p.n_eq, 0);
    std::vector<int> by_key(n);
    std::iota(by_key.begin(), by_key.end(), 0);
    std::stable_sort(by_key.begin(), by_key.end(), [&](int x, int y) { return key[x] < key[y]; });
    size_t seed_pos = 0;
    int n_taken = 0;
    while (n_taken < n) {
        while (taken[by_key[seed_pos]]) ++seed_pos;
        std::vector<int> rows, eqs, cand;
        std::vector<char> is_cand(n, 0);
        auto add_row = [&](int f) {
            taken[f] = 1; ++n_taken; rows.push_back(f);
            for (auto &kc : inc[f]) {
                if (!in_eqs[kc.first]) { in_eqs[kc.first] = 1; eqs.push_back(kc.first); }
                for (int g : rows_of_eq[kc.first]) if (!taken[g] && !is_cand[g]) { is_cand[g] = 1; cand.push_back(g); }
            }
        };
        add_row(by_key[seed_pos]);
        while ((int)rows.size() < rows_per_block && n_taken < n) {
            int best = -1, best_new = 1 << 30;
            size_t w = 0;
            for (size_t i = 0; i < cand.size(); ++i) {
                int g = cand[i];
                if (taken[g]) continue;
                cand[w++] = g;
                int fresh = 0;
                for (auto &kc : inc[g]) fresh += !in_eqs[kc.first];
                if (fresh < best_new || (fresh == best_new && key[g] < key[best])) { best = g; best_new = fresh; }
            }
            cand.resize(w);
            if (best < 0) {                         // component exhausted: continue from the next seed
                while (seed_pos < (size_t)n && taken[by_key[seed_pos]]) ++seed_pos;
                if (seed_pos >= (size_t)n) break;
                best = by_key[seed_pos];
                best_new = 0;
                for (auto &kc : inc[best]) best_new += !in_eqs[kc.first];
            }
            if ((int)eqs.size() + best_new > max_eq_per_block) break;
            add_row(best);
        }
        for (int k : eqs) in_eqs[k] = 0;
        std::sort(eqs.begin(), eqs.end());
        // Colour the block's equations so that no two of one colour touch the same block row: the kernel adds the
        // corner vectors of a colour into the row accumulators without atomics, colours in a fixed order.
        std::vector<int> local_row(n, -1);
        for (size_t i = 0; i < rows.size(); ++i) local_row[rows[i]] = (int)i;
        std::vector<uint32_t> row_colours(rows.size(), 0);
        std::vector<int> colour(eqs.size(), 0);
        int n_colours = 0;
        // pass 1: first-fit gives the number of colours; pass 2: same number, each equation into the emptiest
        // colour its rows allow, so the colours (= the work between two barriers) come out even
        for (int pass = 0; pass < 2; ++pass) {
            std::fill(row_colours.begin(), row_colours.end(), 0u);
            std::vector<int> load(ASM_MAX_COLOURS, 0);
            const int limit = pass == 0 ? ASM_MAX_COLOURS : n_colours;
            for (size_t i = 0; i < eqs.size(); ++i) {
                uint32_t used = 0;
                for (int g : rows_of_eq[eqs[i]]) if (local_row[g] >= 0) used |= row_colours[local_row[g]];
                int c = -1;
                for (int q = 0; q < limit; ++q) {
                    if (used >> q & 1u) continue;
                    if (pass == 0) { c = q; break; }
                    // a colour whose last round of ASM_WARPS_PER_BLOCK equations is still open comes first (its warps
                    // would idle otherwise), then the emptiest
                    auto key = [&](int x) { return (load[x] % ASM_WARPS_PER_BLOCK == 0 ? (1 << 20) : 0) + load[x]; };
                    if (c < 0 || key(q) < key(c)) c = q;
                }
                if (c < 0) {
                    for (int q = limit; q < ASM_MAX_COLOURS && c < 0; ++q) if (!(used >> q & 1u)) c = q;
                    if (c < 0) throw std::runtime_error("assembly plan: more than 32 equation colours in a row block");
                }
                colour[i] = c;
                ++load[c];
                n_colours = std::max(n_colours, c + 1);
                for (int g : rows_of_eq[eqs[i]]) if (local_row[g] >= 0) row_colours[local_row[g]] |= 1u << c;
            }
        }
        // A colour costs ceil(load / warps) rounds (one equation per warp and round, a barrier at its end): empty the
        // last, partly filled round of a colour into the open rounds of the others where the rows allow it, cheapest
        // colour first, until nothing moves.
        {
            constexpr int W = ASM_WARPS_PER_BLOCK;
            std::vector<int> load(ASM_MAX_COLOURS, 0);
            for (int c : colour) ++load[c];
            for (bool moved = true; moved;) {
                moved = false;
                std::vector<int> by_rest;
                for (int c = 0; c < n_colours; ++c) if (load[c] % W) by_rest.push_back(c);
                std::sort(by_rest.begin(), by_rest.end(), [&](int a, int b) { return load[a] % W < load[b] % W; });
                for (int c : by_rest) {
                    // only worth it if the whole rest of the colour finds room
                    std::vector<std::pair<size_t, int>> plan_moves;
                    std::vector<int> room(ASM_MAX_COLOURS, 0);
                    for (int q = 0; q < n_colours; ++q) if (q != c && load[q] % W) room[q] = W - load[q] % W;
                    int need = load[c] % W;
                    std::vector<uint32_t> rc = row_colours;        // tentative
                    for (size_t i = 0; i < eqs.size() && need > 0; ++i) {
                        if (colour[i] != c) continue;
                        uint32_t used = 0;
                        for (int g : rows_of_eq[eqs[i]]) if (local_row[g] >= 0) used |= rc[local_row[g]];
                        used &= ~(1u << c);
                        int best = -1;
                        for (int q = 0; q < n_colours; ++q)
                            if (room[q] > 0 && !(used >> q & 1u) && (best < 0 || room[q] < room[best])) best = q;
                        if (best < 0) continue;
                        plan_moves.push_back({i, best});
                        --room[best];
                        --need;
                        for (int g : rows_of_eq[eqs[i]]) if (local_row[g] >= 0) rc[local_row[g]] = (rc[local_row[g]] & ~(1u << c)) | (1u << best);
                    }
                    if (need > 0) continue;
                    for (auto &mv : plan_moves) { colour[mv.first] = mv.second; --load[c]; ++load[mv.second]; }
                    row_colours.swap(rc);
                    moved = !plan_moves.empty();
                    if (moved) break;
                }
            }
            // colours emptied on the way are closed up
            std::vector<int> remap(ASM_MAX_COLOURS, -1);
            int nc = 0;
            for (int c = 0; c < n_colours; ++c) if (load[c] > 0) remap[c] = nc++;
            for (int &c : colour) c = remap[c];
            n_colours = nc;
        }
        std::vector<int> order(eqs.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return colour[x] < colour[y]; });
        AssemblyBlock blk;
        blk.eq_begin = (int)ap.eq_id.size();
        blk.row_begin = (int)ap.row_perm.size();
        std::vector<int32_t> cptr(ASM_MAX_COLOURS + 1, (int32_t)eqs.size());
        for (size_t i = 0; i < order.size(); ++i) {
            const int k = eqs[order[i]];
            if (i == 0 || colour[order[i]] != colour[order[i - 1]])
                for (int c = (i == 0 ? 0 : colour[order[i - 1]] + 1); c <= colour[order[i]]; ++c) cptr[c] = (int32_t)i;
            ap.eq_id.push_back(k);
            const double *u = &p.tri_u[(size_t)p.eq_tri[k] * 6];
            for (int d = 0; d < 6; ++d) ap.eq_u.push_back((float)u[d]);
            ap.eq_u.push_back(0.f); ap.eq_u.push_back(0.f);          // 8 floats per entry: two aligned float4 loads
            const uint32_t *t = &p.tris[(size_t)p.eq_tri[k] * 3];
            for (int c = 0; c < 3; ++c) {
                const int f = p.vi_to_free[t[c]];
                ap.eq_rows.push_back((int16_t)(f >= 0 ? local_row[f] : -1));
            }
            ap.eq_rows.push_back(0);
        }
        ap.colour_ptr.insert(ap.colour_ptr.end(), cptr.begin(), cptr.end());
        // per-warp walk of the block: the warp's equations of colour 0, a barrier, colour 1, ...
        for (int w = 0; w < ASM_WARPS_PER_BLOCK; ++w) {
            ap.warp_ptr.push_back((int32_t)ap.warp_sched.size());
            for (int c = 0; c < n_colours; ++c) {
                for (int e = cptr[c] + w; e < cptr[c + 1]; e += ASM_WARPS_PER_BLOCK) ap.warp_sched.push_back((int16_t)e);
                ap.warp_sched.push_back(ASM_SCHED_BARRIER);
            }
            ap.warp_sched.push_back(ASM_SCHED_END);
        }
        // row -> incident (equation, corner) pairs of the block (CSR), for the frame-at-a-time gather variant
        {
            std::vector<int> pos(eqs.size());
            for (size_t i = 0; i < order.size(); ++i) pos[order[i]] = (int)i;
            for (int f : rows) {
                ap.row_perm.push_back(p.scratch_row[f]);
                for (auto &kc : inc[f]) {
                    const int sorted = (int)(std::lower_bound(eqs.begin(), eqs.end(), kc.first) - eqs.begin());
                    ap.inc.push_back((uint16_t)(pos[sorted] * 3 + kc.second));
                }
                ap.row_ptr.push_back((int32_t)ap.inc.size());
            }
        }
        blk.eq_end = (int)ap.eq_id.size();
        blk.row_end = (int)ap.row_perm.size();
        blk.n_colours = n_colours;
        ap.max_eq_per_block = std::max(ap.max_eq_per_block, blk.eq_end - blk.eq_begin);
        ap.max_rows_per_block = std::max(ap.max_rows_per_block, blk.row_end - blk.row_begin);
        ap.blocks.push_back(blk);
    }
    ap.warp_ptr.push_back((int32_t)ap.warp_sched.size());
    // frame-tiled compact dgrad: a scale part (6 slots per active equation) and a rotation part (3), each padded
    // to whole 256-row GEMM tiles of the decode kernel; an equation shared by two row blocks is decoded once
    {
        // slots in the order the row blocks walk their equations, so that a block reads (mostly) one contiguous run
        std::vector<int> slot_of_eq(p.n_eq, -1);
        ap.eq_slot.clear();
        ap.slot_eq.clear();
        for (int k : ap.eq_id) {
            if (slot_of_eq[k] < 0) { slot_of_eq[k] = (int)ap.slot_eq.size(); ap.slot_eq.push_back(k); }
            ap.eq_slot.push_back(slot_of_eq[k]);
        }
        const int E = (int)p.active_eq.size();
        ap.compact_s_rows = (6 * E + 255) / 256 * 256;
        ap.compact_stride = ap.compact_s_rows + (3 * E + 255) / 256 * 256;
    }
}

}  // namespace sdfa
