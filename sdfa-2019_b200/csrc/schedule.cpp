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
#include <stdexcept>
#include <cstddef>

namespace sdfa {

namespace {

struct Task {
    int target_slot;
    uint32_t flags;
    float dinv;
    std::vector<TaskEntry> entries;
};

class Emitter {
public:
    explicit Emitter(SolveProgram &prog) : prog_(prog) { open_stage(); }

    void rows(std::vector<Task> &tasks, bool sync_after) {
        // longest first so the round-robin warp assignment of the kernel is balanced
        std::stable_sort(tasks.begin(), tasks.end(),
                         [](const Task &a, const Task &b) { return a.entries.size() > b.entries.size(); });
        size_t i = 0;
        while (i < tasks.size()) {
            // how many tasks fit into what is left of this stage
            size_t room = STAGE_BYTES - cur_.size();
            size_t n = 0, bytes = sizeof(OpHeader);
            while (i + n < tasks.size()) {
                size_t table = ((n + 1) * 4 + 15) / 16 * 16;
                size_t body = task_bytes(tasks[i + n]);
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
                const Task &tk = tasks[i + t];
                uint32_t off = (uint32_t)cur_.size();
                std::memcpy(&cur_[table_at + t * 4], &off, 4);
                size_t ne = tk.entries.size() + (tk.entries.size() & 1);   // pad to even -> 16 B multiple
                TaskHeader th{(uint32_t)tk.target_slot * SLOT_BYTES, (uint32_t)ne | tk.flags, tk.dinv, 0};
                cur_.resize(off + sizeof(TaskHeader) + ne * sizeof(TaskEntry), 0);
                std::memcpy(&cur_[off], &th, sizeof(th));
                if (!tk.entries.empty())
                    std::memcpy(&cur_[off + sizeof(TaskHeader)], tk.entries.data(), tk.entries.size() * sizeof(TaskEntry));
                prog_.n_entries += (long long)tk.entries.size();
            }
            uint32_t next = (uint32_t)cur_.size();
            std::memcpy(&cur_[op_at + offsetof(OpHeader, c)], &next, 4);
            ++n_ops_;
            i += n;
        }
        if (tasks.empty() && sync_after) { /* nothing to wait for */ }
    }

    void rowspan(OpType type, int first_row, const std::vector<uint32_t> &table, uint16_t flags) {
        size_t need = sizeof(OpHeader) + (table.size() * 4 + 15) / 16 * 16;
        if (cur_.size() + need > STAGE_BYTES) { close_stage(); open_stage(); }
        size_t op_at = cur_.size();
        OpHeader h{(uint16_t)type, flags, (uint32_t)first_row, (uint32_t)table.size(), (uint32_t)(op_at + sizeof(OpHeader))};
        cur_.resize(op_at + need, 0);
        std::memcpy(&cur_[op_at], &h, sizeof(h));
        std::memcpy(&cur_[op_at + sizeof(OpHeader)], table.data(), table.size() * 4);
        ++n_ops_;
    }

    void finish() {
        close_stage();
        prog_.stage_off.push_back((uint32_t)prog_.bytes.size());
    }

private:
    static size_t task_bytes(const Task &t) {
        size_t ne = t.entries.size() + (t.entries.size() & 1);
        return sizeof(TaskHeader) + ne * sizeof(TaskEntry);
    }
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

class SlotPool {
public:
    int alloc() {
        int s;
        if (!free_.empty()) { s = free_.back(); free_.pop_back(); }
        else s = next_++;
        ++live_;
        peak_ = std::max(peak_, next_);
        return s;
    }
    void release(int s) { free_.push_back(s); --live_; }
    int peak() const { return peak_; }
    // hand out the lowest free ids first: keeps pieces mostly contiguous
    void sort_free() { std::sort(free_.begin(), free_.end(), std::greater<int>()); }
private:
    std::vector<int> free_;
    int next_ = 0, live_ = 0, peak_ = 0;
};

}  // namespace

void build_solve_program(HostPlan &p, int piece_cap) {
    const int n = p.n_free;
    SolveProgram &prog = p.prog;
    prog = SolveProgram();
    Emitter em(prog);

    // row-wise view of the strict lower triangle
    std::vector<int> rptr(n + 1, 0);
    for (int j = 0; j < n; ++j)
        for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) rptr[p.l_rowidx[q] + 1]++;
    for (int i = 0; i < n; ++i) rptr[i + 1] += rptr[i];
    std::vector<int> rcol(rptr[n]);
    std::vector<float> rval(rptr[n]);
    {
        std::vector<int> fill(rptr.begin(), rptr.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) {
                int i = p.l_rowidx[q];
                rcol[fill[i]] = j;                      // ascending j because columns are visited in order
                rval[fill[i]++] = (float)p.l_val[q];
            }
    }
    std::vector<float> dinv(n);
    for (int j = 0; j < n; ++j) dinv[j] = (float)(1.0 / p.l_val[p.l_colptr[j]]);

    std::vector<std::pair<int, int>> pieces;
    for (int a = 0; a < n; a += piece_cap) pieces.push_back({a, std::min(n, a + piece_cap)});
    prog.n_pieces = (int)pieces.size();

    std::vector<int> slot_of(n, -1), lvl(n, 0);
    // ---------------------------------------------------------------- forward sweep
    {
        SlotPool pool;
        for (auto [a, b] : pieces) {
            std::vector<uint32_t> table;
            for (int i = a; i < b; ++i) {
                uint32_t add = 0;
                if (slot_of[i] < 0) slot_of[i] = pool.alloc();
                else add = LOAD_ADD_BIT;                 // holds partial sums pushed by earlier pieces
                table.push_back((uint32_t)(slot_of[i] * SLOT_WORDS) | add);
            }
            em.rowspan(OP_LOAD, a, table, OPF_SYNC_AFTER);
            std::vector<std::vector<Task>> levels;
            auto put = [&](int level, Task &&t) {
                if ((int)levels.size() <= level) levels.resize(level + 1);
                levels[level].push_back(std::move(t));
            };
            for (int i = a; i < b; ++i) {
                Task t{slot_of[i], TASK_FINAL, dinv[i], {}};
                int l = 0;
                for (int q = rptr[i]; q < rptr[i + 1]; ++q) {
                    int j = rcol[q];
                    if (j < a) continue;                 // already pushed into the slot by j's piece
                    l = std::max(l, lvl[j] + 1);
                    t.entries.push_back({rval[q], (uint32_t)(slot_of[j] * SLOT_BYTES)});
                }
                lvl[i] = l;
                put(l, std::move(t));
            }
            // contributions of this piece to rows above it
            std::map<int, Task> ext;
            std::map<int, int> ext_lvl;
            for (int j = a; j < b; ++j)
                for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) {
                    int i = p.l_rowidx[q];
                    if (i < b) continue;
                    auto it = ext.find(i);
                    if (it == ext.end()) {
                        uint32_t fl = 0;
                        if (slot_of[i] < 0) { slot_of[i] = pool.alloc(); fl = TASK_OVERWRITE; }
                        it = ext.emplace(i, Task{slot_of[i], fl, 1.f, {}}).first;
                        ext_lvl[i] = 0;
                    }
                    it->second.entries.push_back({(float)p.l_val[q], (uint32_t)(slot_of[j] * SLOT_BYTES)});
                    ext_lvl[i] = std::max(ext_lvl[i], lvl[j] + 1);
                }
            for (auto &kv : ext) put(ext_lvl[kv.first], std::move(kv.second));
            for (auto &lv : levels) { em.rows(lv, true); prog.n_steps_fwd++; }
            for (auto &w : table) w &= ~LOAD_ADD_BIT;
            em.rowspan(OP_STORE_Y, a, table, OPF_SYNC_AFTER);
            for (int i = a; i < b; ++i) { pool.release(slot_of[i]); slot_of[i] = -1; }
            pool.sort_free();
        }
        prog.n_slots = std::max(prog.n_slots, pool.peak());
    }
    // ---------------------------------------------------------------- backward sweep
    {
        SlotPool pool;
        std::fill(slot_of.begin(), slot_of.end(), -1);
        // row i must stay resident until the piece holding the smallest column of its row pattern is done
        std::vector<int> piece_of(n);
        for (int pi = 0; pi < (int)pieces.size(); ++pi)
            for (int i = pieces[pi].first; i < pieces[pi].second; ++i) piece_of[i] = pi;
        std::vector<std::vector<int>> release_after(pieces.size());
        for (int i = 0; i < n; ++i) {
            int last = (rptr[i + 1] > rptr[i]) ? piece_of[rcol[rptr[i]]] : piece_of[i];
            release_after[last].push_back(i);
        }
        for (int pi = (int)pieces.size() - 1; pi >= 0; --pi) {
            auto [a, b] = pieces[pi];
            std::vector<uint32_t> table;
            for (int i = a; i < b; ++i) {
                slot_of[i] = pool.alloc();
                table.push_back((uint32_t)(slot_of[i] * SLOT_WORDS));
            }
            em.rowspan(OP_LOAD, a, table, OPF_SYNC_AFTER);
            std::vector<std::vector<Task>> levels;
            for (int j = b - 1; j >= a; --j) {
                Task t{slot_of[j], TASK_FINAL, dinv[j], {}};
                int l = 0;
                for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) {
                    int i = p.l_rowidx[q];
                    assert(slot_of[i] >= 0);
                    if (i < b) l = std::max(l, lvl[i] + 1);
                    t.entries.push_back({(float)p.l_val[q], (uint32_t)(slot_of[i] * SLOT_BYTES)});
                }
                lvl[j] = l;
                if ((int)levels.size() <= l) levels.resize(l + 1);
                levels[l].push_back(std::move(t));
            }
            for (auto &lv : levels) { em.rows(lv, true); prog.n_steps_bwd++; }
            em.rowspan(OP_STORE_X, a, table, OPF_SYNC_AFTER);
            for (int i : release_after[pi]) { pool.release(slot_of[i]); slot_of[i] = -1; }
            pool.sort_free();
        }
        prog.n_slots = std::max(prog.n_slots, pool.peak());
    }
    em.finish();
}

// ------------------------------------------------------------------------------------------
void build_assembly_plan(HostPlan &p, int rows_per_block) {
    AssemblyPlan &ap = p.asmplan;
    ap = AssemblyPlan();
    const int n = p.n_free;
    // incidences per free column: (equation block, corner)
    std::vector<std::vector<std::pair<int, int>>> inc(n);
    for (int k : p.active_eq) {
        const uint32_t *t = &p.tris[(size_t)p.eq_tri[k] * 3];
        for (int c = 0; c < 3; ++c) {
            int f = p.vi_to_free[t[c]];
            if (f >= 0) inc[f].push_back({k, c});
        }
    }
    // group rows whose equations sit close together in the dgrad row: sort by smallest incident equation
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::vector<int> key(n, 0x7fffffff);
    for (int f = 0; f < n; ++f)
        for (auto &kc : inc[f]) key[f] = std::min(key[f], kc.first);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
    ap.row_ptr.push_back(0);
    for (int start = 0; start < n; start += rows_per_block) {
        int stop = std::min(n, start + rows_per_block);
        AssemblyBlock blk;
        blk.eq_begin = (int)ap.eq_id.size();
        blk.row_begin = (int)ap.row_perm.size();
        std::vector<int> eqs;
        for (int r = start; r < stop; ++r)
            for (auto &kc : inc[order[r]]) eqs.push_back(kc.first);
        std::sort(eqs.begin(), eqs.end());
        eqs.erase(std::unique(eqs.begin(), eqs.end()), eqs.end());
        for (int k : eqs) {
            ap.eq_id.push_back(k);
            const double *u = &p.tri_u[(size_t)p.eq_tri[k] * 6];
            for (int d = 0; d < 6; ++d) ap.eq_u.push_back((float)u[d]);
        }
        for (int r = start; r < stop; ++r) {
            int f = order[r];
            ap.row_perm.push_back(p.iperm[f]);
            for (auto &kc : inc[f]) {
                int local = (int)(std::lower_bound(eqs.begin(), eqs.end(), kc.first) - eqs.begin());
                ap.inc.push_back((uint16_t)(local * 3 + kc.second));
            }
            ap.row_ptr.push_back((int32_t)ap.inc.size());
        }
        blk.eq_end = (int)ap.eq_id.size();
        blk.row_end = (int)ap.row_perm.size();
        ap.max_eq_per_block = std::max(ap.max_eq_per_block, blk.eq_end - blk.eq_begin);
        ap.max_rows_per_block = std::max(ap.max_rows_per_block, blk.row_end - blk.row_begin);
        ap.blocks.push_back(blk);
    }
}

}  // namespace sdfa
