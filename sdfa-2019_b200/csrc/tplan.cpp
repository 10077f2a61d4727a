// tplan.cpp -- host planner of the tensor-core solve (see tplan.hpp for the scheme and the formats).
//
// Replaces, together with solve_tc.cu, the per-frame `solver_.solve(...)` of the reference
// (deformation/cpp/src/deform_triangle_impl.hpp:286-292; Eigen::SparseLU set up at :122-139) for templates
// whose system fits: the factorisation below is a block Cholesky over a nested-dissection tree, computed
// once in fp64, whose blocks are folded into explicit products so that the GPU only multiplies.
#include "plan.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>

namespace sdfa {

namespace {

inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

inline float rn_tf32(float x) {          // round to nearest-even TF32 (10 explicit mantissa bits), low 13 bits zero
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

struct Coupling {
    int anc;                      // ancestor node index
    int r0, r1;                   // coupled local rows of the ancestor lie in [r0, r1)
    std::vector<double> G;        // |anc| x |J| row-major: F_aJ F_JJ^-1 (zero rows where not coupled)
};

struct TNode {
    std::vector<int> rows;        // free columns of this node, in scratch order
    int parent = -1;
    std::vector<int> children;
    std::vector<int> anc;         // ancestors, nearest first
    int row0 = 0;                 // first scratch row
    int acc_off = 0, xs_off = 0;  // tensor-memory stack offsets (forward accumulators / backward solutions)
    std::vector<double> P;        // |J| x |J|: F_JJ^-1
    std::vector<Coupling> coup;
};

struct Builder {
    HostPlan &p;
    TensorPlan &t;
    int n;
    std::vector<std::vector<int>> adj;
    std::vector<TNode> nodes;
    int leaf_max;
    int merge_cap = TS_MAX_NODE;          // largest separator node that folding may create
    int used_slots = 0, used_ring = 0;    // pipelining depth the tensor-memory budget allowed (emit)
    std::string err;

    Builder(HostPlan &hp, int lm) : p(hp), t(hp.tplan), n(hp.n_free), leaf_max(lm) {
        if (const char *e = std::getenv("SDFA_TS_STREAMS")) n_streams = std::atoi(e) == 1 ? 1 : 2;
        if (const char *e = std::getenv("SDFA_TS_EARLY")) early_signals = std::atoi(e) != 0;
    }

    const float *pos(int f) const { return &p.verts[(size_t)p.free_to_vi[f] * 3]; }

    static int widest_axis(const std::vector<int> &idx, const std::function<const float *(int)> &pos) {
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int f : idx)
            for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], pos(f)[d]); hi[d] = std::max(hi[d], pos(f)[d]); }
        int ax = 0;
        for (int d = 1; d < 3; ++d) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
        return ax;
    }

    // geometric nested dissection: split at the median of the widest axis, the separator is the smaller of
    // the two one-sided boundaries of the cut
    int bisect(std::vector<int> idx, int parent) {
        const int me = (int)nodes.size();
        nodes.emplace_back();
        nodes[me].parent = parent;
        auto P = [this](int f) { return pos(f); };
        if ((int)idx.size() <= leaf_max) {
            nodes[me].rows = idx;
            return me;
        }
        const int ax = widest_axis(idx, P);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return pos(a)[ax] < pos(b)[ax]; });
        // cut so that both sides can be tiled by the fewest leaves: k = leaves needed for this part, the low side
        // gets floor(k/2) of them (a plain median cut turns a 130-row part into four leaves instead of three)
        const int k_leaves = ((int)idx.size() + leaf_max - 1) / leaf_max;
        const int half = std::getenv("SDFA_TS_MEDIAN") ? (int)idx.size() / 2
                                                        : (int)((long long)idx.size() * (k_leaves / 2) / k_leaves);
        std::vector<int> side(n, -1);                       // -1 outside, 0 / 1 the two sides
        for (int i = 0; i < (int)idx.size(); ++i) side[idx[i]] = i < half ? 0 : 1;
        std::vector<int> bnd[2];
        for (int f : idx)
            for (int g : adj[f])
                if (side[g] >= 0 && side[g] != side[f]) { bnd[side[f]].push_back(f); break; }
        const std::vector<int> &sep = bnd[0].size() <= bnd[1].size() ? bnd[0] : bnd[1];
        std::vector<char> in_sep(n, 0);
        for (int f : sep) in_sep[f] = 1;
        std::vector<int> part[2], s(sep);
        for (int f : idx) if (!in_sep[f]) part[side[f]].push_back(f);
        if (part[0].empty() || part[1].empty()) {           // degenerate cut: keep as one block if it fits
            nodes[me].rows = idx;
            if ((int)idx.size() > TS_MAX_NODE) err = "cannot split a block of " + std::to_string(idx.size()) + " rows";
            return me;
        }
        if (!s.empty()) {                                   // order the separator along its own widest axis
            const int sax = widest_axis(s, P);
            std::stable_sort(s.begin(), s.end(), [&](int a, int b) { return pos(a)[sax] < pos(b)[sax]; });
        }
        nodes[me].rows = s;
        for (int k = 0; k < 2; ++k) {
            const int c = bisect(part[k], me);
            nodes[me].children.push_back(c);
        }
        return me;
    }

    // Every separator is a serial hand-off between the two instruction streams in both sweeps, and bisection leaves
    // many tiny ones near the leaves: fold a separator into its parent separator while the union still fits a tree
    // node.  The union is eliminated after everything below either of them, so the block factorisation stays exact;
    // the folded node's children move up.
    void coarsen(int v) {
        for (size_t i = 0; i < nodes[v].children.size(); ++i) coarsen(nodes[v].children[i]);
        if (nodes[v].children.empty()) return;
        const int cap = std::min(TS_MAX_NODE, merge_cap);
        for (bool again = true; again;) {
            again = false;
            int best = -1;
            for (int c : nodes[v].children)
                if (!nodes[c].children.empty() && (int)(nodes[v].rows.size() + nodes[c].rows.size()) <= cap &&
                    (best < 0 || nodes[c].rows.size() < nodes[best].rows.size())) best = c;
            if (best < 0) break;
            TNode &p = nodes[v], &c = nodes[best];
            p.rows.insert(p.rows.end(), c.rows.begin(), c.rows.end());
            p.children.erase(std::find(p.children.begin(), p.children.end(), best));
            for (int g : c.children) { nodes[g].parent = v; p.children.push_back(g); }
            c.rows.clear();
            c.children.clear();
            c.parent = -2;                                        // dead
            again = true;
        }
    }

    void postorder(int v, std::vector<int> &out) {
        for (int c : nodes[v].children) postorder(c, out);
        out.push_back(v);
    }

    // in-place inverse of a symmetric positive definite k x k matrix through its Cholesky factor;
    // returns false when a pivot is not safely positive
    static bool spd_inverse(std::vector<double> &a, int k, double pivot_floor) {
        std::vector<double> l((size_t)k * k, 0.0);
        for (int j = 0; j < k; ++j) {
            double d = a[(size_t)j * k + j];
            for (int q = 0; q < j; ++q) d -= l[(size_t)j * k + q] * l[(size_t)j * k + q];
            if (!(d > pivot_floor)) return false;
            const double ljj = std::sqrt(d);
            l[(size_t)j * k + j] = ljj;
            for (int i = j + 1; i < k; ++i) {
                double s = a[(size_t)i * k + j];
                for (int q = 0; q < j; ++q) s -= l[(size_t)i * k + q] * l[(size_t)j * k + q];
                l[(size_t)i * k + j] = s / ljj;
            }
        }
        // W = L^-1 (lower), then A^-1 = W^T W
        std::vector<double> w((size_t)k * k, 0.0);
        for (int c = 0; c < k; ++c) {
            w[(size_t)c * k + c] = 1.0 / l[(size_t)c * k + c];
            for (int i = c + 1; i < k; ++i) {
                double s = 0.0;
                for (int q = c; q < i; ++q) s -= l[(size_t)i * k + q] * w[(size_t)q * k + c];
                w[(size_t)i * k + c] = s / l[(size_t)i * k + i];
            }
        }
        for (int i = 0; i < k; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = 0.0;
                for (int q = i; q < k; ++q) s += w[(size_t)q * k + i] * w[(size_t)q * k + j];
                a[(size_t)i * k + j] = a[(size_t)j * k + i] = s;
            }
        return true;
    }

    bool factor(const std::vector<int> &post) {
        std::vector<double> F((size_t)n * n, 0.0);
        double max_diag = 0.0;
        for (int c = 0; c < n; ++c)
            for (int q = p.m_colptr[c]; q < p.m_colptr[c + 1]; ++q) {
                const int r = p.m_rowidx[q];
                F[(size_t)r * n + c] = F[(size_t)c * n + r] = p.m_val[q];
                if (r == c) max_diag = std::max(max_diag, p.m_val[q]);
            }
        const double floor = 1e-9 * max_diag;               // the SIMT path pins such pivots (unconstrained templates)
        for (int v : post) {
            TNode &nd = nodes[v];
            const int k = (int)nd.rows.size();
            if (k == 0) continue;
            nd.P.assign((size_t)k * k, 0.0);
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) nd.P[(size_t)i * k + j] = F[(size_t)nd.rows[i] * n + nd.rows[j]];
            if (!spd_inverse(nd.P, k, floor)) { err = "pivot below 1e-9 x max diagonal (singular / unconstrained system)"; return false; }
            std::vector<int> brow;                           // coupled ancestor rows (free columns)
            std::vector<double> gb, fb;                      // G and F restricted to them, |B| x k
            for (int a : nd.anc) {
                const TNode &an = nodes[a];
                const int m = (int)an.rows.size();
                Coupling c;
                c.anc = a; c.r0 = m; c.r1 = 0;
                std::vector<double> faj((size_t)m * k);
                for (int i = 0; i < m; ++i) {
                    bool nz = false;
                    for (int j = 0; j < k; ++j) {
                        const double x = F[(size_t)an.rows[i] * n + nd.rows[j]];
                        faj[(size_t)i * k + j] = x;
                        nz |= (x != 0.0);
                    }
                    if (nz) { c.r0 = std::min(c.r0, i); c.r1 = std::max(c.r1, i + 1); }
                }
                if (c.r1 <= c.r0) continue;
                c.G.assign((size_t)m * k, 0.0);
                for (int i = c.r0; i < c.r1; ++i) {
                    bool nz = false;
                    for (int j = 0; j < k; ++j) nz |= (faj[(size_t)i * k + j] != 0.0);
                    if (!nz) continue;
                    for (int j = 0; j < k; ++j) {
                        double s = 0.0;
                        for (int q = 0; q < k; ++q) s += faj[(size_t)i * k + q] * nd.P[(size_t)q * k + j];
                        c.G[(size_t)i * k + j] = s;
                    }
                    brow.push_back(an.rows[i]);
                    gb.insert(gb.end(), c.G.begin() + (size_t)i * k, c.G.begin() + (size_t)(i + 1) * k);
                    fb.insert(fb.end(), faj.begin() + (size_t)i * k, faj.begin() + (size_t)(i + 1) * k);
                }
                nd.coup.push_back(std::move(c));
            }
            const int nb = (int)brow.size();                 // Schur update F_BB -= G_B F_BJ^T
            for (int i = 0; i < nb; ++i)
                for (int j = 0; j < nb; ++j) {
                    double s = 0.0;
                    for (int q = 0; q < k; ++q) s += gb[(size_t)i * k + q] * fb[(size_t)j * k + q];
                    F[(size_t)brow[i] * n + brow[j]] -= s;
                }
        }
        return true;
    }

    // ---------------------------------------------------------------- program emission
    // Streams: 0 / 1 = the EPI streams, 2 = the MMA stream.  EPI ops share one index space (t.epi), MMA ops
    // theirs (t.mma).  Per tensor-memory column the last op of every stream that wrote / read it; an op waits
    // for the latest conflicting op of each OTHER stream (read after write, write after read / write).
    static constexpr int SM = 2;
    int n_streams = 2;
    std::vector<int> epi_wait_mma_op, epi_wait_epi_op;      // per EPI op: op index to wait for in the MMA / other EPI stream, -1 none
    std::vector<std::array<int, 2>> mma_wait_epi_op;        // per MMA op: op index per EPI stream
    std::vector<char> epi_signals, epi_signals_read, mma_commits;   // per op: someone waits for its end / for its source read
    std::vector<char> epi_wait_epi_early;                   // per EPI op: the EPI wait is for the other op's read event
    std::vector<std::array<char, 2>> mma_wait_early;        // per MMA op and EPI stream: likewise
    int covered_read[3][2];                                 // [s][q]: latest op of EPI stream q whose source read stream s is known to be after
    bool early_signals = true;
    int last_write[3][TS_TMEM_COLS], last_read[3][TS_TMEM_COLS];
    // every stream runs in order, so having waited for op k of a stream implies all its earlier ops; and a wait
    // is transitive: the op waited for had itself waited (covered[][] at that point) -- skip what is implied
    int covered[3][3];                                      // covered[s][o]: latest op of stream o that stream s is known to be after
    std::vector<std::array<int, 3>> epi_cov, mma_cov;       // snapshot of the op's stream's covered[] row when the op ends
    uint32_t chunk_fill = 0;

    struct Range { int c0, c1; };

    void absorb(int s, int o, int k) {                      // stream s now waits for op k of stream o
        covered[s][o] = std::max(covered[s][o], k);
        const std::array<int, 3> &snap = o == SM ? mma_cov[k] : epi_cov[k];
        for (int q = 0; q < 3; ++q) if (q != s) covered[s][q] = std::max(covered[s][q], snap[q]);
    }

    // wait of stream s for op `need` of EPI stream q; `read_only`: only that op's tensor-memory READ conflicts.
    // Returns -1 when implied by earlier waits, else `need`, with early = the read event suffices.
    int epi_need(int s, int q, int need_w, int need_r, bool &early) {
        early = false;
        if (!early_signals) { need_w = std::max(need_w, need_r); need_r = -1; }
        if (need_w >= need_r) {
            if (need_w <= covered[s][q]) return -1;
            absorb(s, q, need_w);
            covered_read[s][q] = std::max(covered_read[s][q], need_w);
            return need_w;
        }
        if (need_r <= covered[s][q] || need_r <= covered_read[s][q]) return -1;
        covered_read[s][q] = need_r;
        if (need_r > 0) {                                   // everything before it in its stream is complete
            // ops of stream q before need_r: the latest one
            int prev = -1;
            for (int e = need_r - 1; e >= 0; --e) if (t.epi[e].stream == q) { prev = e; break; }
            if (prev > covered[s][q]) absorb(s, q, prev);
        }
        // the op itself has started: what it waited for is implied
        for (int r = 0; r < 3; ++r) if (r != s && r != q) covered[s][r] = std::max(covered[s][r], epi_cov[need_r][r]);
        early = true;
        return need_r;
    }

    void add_epi(EpiOp op, int s, std::vector<Range> reads, std::vector<Range> writes) {
        if (n_streams < 2) s = 0;
        const int e = (int)t.epi.size(), o = 1 - s;
        int need_m = -1, need_ew = -1, need_er = -1;
        for (auto r : reads) for (int c = r.c0; c < r.c1; ++c) {
            need_m = std::max(need_m, last_write[SM][c]);
            need_ew = std::max(need_ew, last_write[o][c]);
        }
        for (auto r : writes) for (int c = r.c0; c < r.c1; ++c) {
            need_m = std::max(need_m, std::max(last_write[SM][c], last_read[SM][c]));
            need_ew = std::max(need_ew, last_write[o][c]);
            need_er = std::max(need_er, last_read[o][c]);
        }
        for (auto r : reads) for (int c = r.c0; c < r.c1; ++c) last_read[s][c] = e;
        for (auto r : writes) for (int c = r.c0; c < r.c1; ++c) last_write[s][c] = e;
        if (need_m <= covered[s][SM]) need_m = -1;
        else absorb(s, SM, need_m);
        bool early = false;
        const int need_e = epi_need(s, o, need_ew, need_er, early);
        if (need_m >= 0) mma_commits[need_m] = 1;
        if (need_e >= 0) (early ? epi_signals_read : epi_signals)[need_e] = 1;
        epi_wait_mma_op.push_back(need_m);
        epi_wait_epi_op.push_back(need_e);
        epi_wait_epi_early.push_back(early);
        epi_signals.push_back(0);
        epi_signals_read.push_back(0);
        covered[s][s] = e;
        epi_cov.push_back({covered[s][0], covered[s][1], covered[s][2]});
        op.wait_mma = op.signal_epi = op.wait_epi = op.signal_read = -1;
        op.stream = (uint16_t)s;
        op.ring_seq = (op.flags & (EPI_ADD_GLOBAL | EPI_STORE_GLOBAL)) ? (uint16_t)t.n_ring_ops++ : (uint16_t)0;
        t.epi.push_back(op);
    }

    // Bt(nn, kk) = element of the [N x K] tile (row nn of the product's output, kk of its source)
    void add_mma(MmaOp op, const std::function<double(int, int)> &bt) {
        const int m = (int)t.mma.size();
        const int K = op.k8 * 8, N = op.n;
        std::array<int, 2> need = {-1, -1};
        std::array<char, 2> early = {0, 0};
        for (int q = 0; q < 2; ++q) {
            int need_w = -1, need_r = -1;
            for (int c = 0; c < K; ++c) need_w = std::max(need_w, std::max(last_write[q][op.a_hi_col + c], last_write[q][op.a_lo_col + c]));
            for (int c = 0; c < N; ++c) {
                need_w = std::max(need_w, last_write[q][op.d_col + c]);
                need_r = std::max(need_r, last_read[q][op.d_col + c]);
            }
            bool ea = false;
            need[q] = epi_need(SM, q, need_w, need_r, ea);
            early[q] = ea;
        }
        for (int c = 0; c < K; ++c) last_read[SM][op.a_hi_col + c] = last_read[SM][op.a_lo_col + c] = m;
        for (int c = 0; c < N; ++c) last_write[SM][op.d_col + c] = m;
        for (int q = 0; q < 2; ++q) if (need[q] >= 0) (early[q] ? epi_signals_read : epi_signals)[need[q]] = 1;
        mma_wait_early.push_back(early);
        mma_wait_epi_op.push_back(need);
        mma_commits.push_back(0);
        covered[SM][SM] = m;
        mma_cov.push_back({covered[SM][0], covered[SM][1], covered[SM][2]});
        // tile images: K-blocks of 32, each an [N x 32] K-major SWIZZLE_128B image; hi then lo
        const int kb = (K + 31) / 32;
        const uint32_t half = (uint32_t)kb * N * 128, bytes = 2 * half;
        if (t.chunk_off.empty()) { t.chunk_off.push_back(0); chunk_fill = 0; op.flags |= MMA_CHUNK_FIRST; }
        else if (chunk_fill + bytes > (uint32_t)TS_STAGE_BYTES) {
            t.mma.back().flags |= MMA_CHUNK_LAST;
            t.chunk_off.push_back((uint32_t)t.matrix.size());
            chunk_fill = 0;
            op.flags |= MMA_CHUNK_FIRST;
        }
        op.b_hi_off = chunk_fill;
        op.b_lo_off = chunk_fill + half;
        chunk_fill += bytes;
        const size_t base = t.matrix.size();
        t.matrix.resize(base + bytes, 0);
        float *hi = reinterpret_cast<float *>(t.matrix.data() + base), *lo = reinterpret_cast<float *>(t.matrix.data() + base + half);
        for (int nn = 0; nn < N; ++nn)
            for (int kk = 0; kk < K; ++kk) {
                const double v = bt(nn, kk);
                if (v == 0.0) continue;
                const float h = rn_tf32((float)v), l = rn_tf32((float)(v - (double)h));
                const int at = (kk / 32) * (N * 32) + ts_swz(nn, kk % 32);
                hi[at] = h;
                lo[at] = l;
            }
        op.wait_epi = op.wait_epi2 = op.commit_mma = -1;
        t.mma.push_back(op);
        t.nk_products += (long long)N * K;
    }

    bool emit(const std::vector<int> &post) {
        for (int q = 0; q < 3; ++q) {
            std::fill(last_write[q], last_write[q] + TS_TMEM_COLS, -1);
            std::fill(last_read[q], last_read[q] + TS_TMEM_COLS, -1);
            for (int r = 0; r < 3; ++r) covered[q][r] = -1;
            covered_read[q][0] = covered_read[q][1] = -1;
        }
        t.n_streams = n_streams;
        std::vector<int> steps;
        int kmax = 0, nmax = 0, acc_cols = 0, xs_cols = 0;
        for (int v : post) {
            const int k = (int)nodes[v].rows.size();
            if (k == 0) continue;
            steps.push_back(v);
            kmax = std::max(kmax, pad_to(k, 8));
            nmax = std::max(nmax, pad_to(k, 16));
            if (!nodes[v].children.empty()) {
                acc_cols = std::max(acc_cols, nodes[v].acc_off + pad_to(k, 16));
                xs_cols = std::max(xs_cols, nodes[v].xs_off + pad_to(k, 8));
            }
        }
        // ---- tensor-memory maps
        const int s0 = acc_cols;
        int n_slots = 3;                                    // source slots: more of them let the loads run further ahead
        if (const char *ns = std::getenv("SDFA_TS_SLOTS")) n_slots = std::max(1, std::min(4, std::atoi(ns)));
        while (n_slots > 1 && pad_to(s0 + n_slots * 2 * kmax, 16) + nmax > TS_TMEM_COLS) --n_slots;
        const int du = pad_to(s0 + n_slots * 2 * kmax, 16);
        t.tmem_fwd = du + nmax;
        if (t.tmem_fwd > TS_TMEM_COLS) { err = "forward sweep needs " + std::to_string(t.tmem_fwd) + " tensor-memory columns"; return false; }
        const int dx0 = pad_to(2 * xs_cols, 16);
        const int ring = std::min(3, (TS_TMEM_COLS - dx0) / std::max(nmax, 1));
        if (ring < 1) { err = "backward sweep needs " + std::to_string(dx0 + nmax) + " tensor-memory columns"; return false; }
        t.tmem_bwd = dx0 + ring * nmax;
        used_slots = n_slots;
        used_ring = ring;
        auto slot_hi = [&](int k) { return s0 + k * 2 * kmax; };
        auto slot_lo = [&](int k) { return s0 + k * 2 * kmax + kmax; };

        // ---- forward sweep
        // the accumulator stack starts at zero; a separator's columns are zeroed again when it is read
        for (int c0 = 0; c0 < acc_cols; c0 += TS_MAX_NODE) {
            EpiOp op{};
            op.flags = EPI_ST_RAW;
            op.n_chunks = (uint16_t)(std::min(TS_MAX_NODE, acc_cols - c0) / 8);
            op.hi_col = (uint16_t)c0;
            add_epi(op, 0, {}, {{c0, c0 + op.n_chunks * 8}});
        }
        auto emit_prep = [&](int i) {
            const TNode &nd = nodes[steps[i]];
            const int k = (int)nd.rows.size(), k8 = pad_to(k, 8), sl = i % n_slots;
            EpiOp op{};
            op.n_chunks = (uint16_t)(k8 / 8);
            op.n_valid = (uint16_t)k;
            op.hi_col = (uint16_t)slot_hi(sl);
            op.lo_col = (uint16_t)slot_lo(sl);
            std::vector<Range> rd;
            if (nd.children.empty()) {
                op.flags = EPI_ADD_GLOBAL | EPI_ST_SPLIT;
                op.row_in = (uint32_t)nd.row0;
            } else {
                op.flags = EPI_FROM_TMEM | EPI_ADD_GLOBAL | EPI_ST_SPLIT | EPI_ZERO_SRC;
                op.src_col = (uint16_t)nd.acc_off;
                op.row_in = (uint32_t)nd.row0;
                rd.push_back({nd.acc_off, nd.acc_off + k8});
            }
            std::vector<Range> wr = {{slot_hi(sl), slot_hi(sl) + k8}, {slot_lo(sl), slot_lo(sl) + k8}};
            if (!nd.children.empty()) wr.push_back({nd.acc_off, nd.acc_off + k8});
            add_epi(op, 0, rd, wr);                           // the forward sweep's loads: EPI stream 0
        };
        const int ns = (int)steps.size();
        if (ns == 0) { err = "empty system"; return false; }
        emit_prep(0);
        for (int i = 0; i < ns; ++i) {
            const TNode &nd = nodes[steps[i]];
            const int k = (int)nd.rows.size(), k8 = pad_to(k, 8), sl = i % n_slots;
            for (const Coupling &c : nd.coup) {
                const TNode &an = nodes[c.anc];
                const int m = (int)an.rows.size();
                const int w0 = c.r0 / 16 * 16, w1 = pad_to(c.r1, 16);
                MmaOp op{};
                op.d_col = (uint16_t)(an.acc_off + w0);
                op.a_hi_col = (uint16_t)slot_hi(sl);
                op.a_lo_col = (uint16_t)slot_lo(sl);
                op.n = (uint16_t)(w1 - w0);
                op.k8 = (uint16_t)(k8 / 8);
                op.flags = MMA_ACCUMULATE;
                add_mma(op, [&](int nn, int kk) { return (w0 + nn < m && kk < k) ? -c.G[(size_t)(w0 + nn) * k + kk] : 0.0; });
            }
            {
                MmaOp op{};
                op.d_col = (uint16_t)du;
                op.a_hi_col = (uint16_t)slot_hi(sl);
                op.a_lo_col = (uint16_t)slot_lo(sl);
                op.n = (uint16_t)pad_to(k, 16);
                op.k8 = (uint16_t)(k8 / 8);
                add_mma(op, [&](int nn, int kk) { return (nn < k && kk < k) ? nd.P[(size_t)nn * k + kk] : 0.0; });
            }
            if (i + 1 < ns) emit_prep(i + 1);
            EpiOp op{};
            op.flags = EPI_FROM_TMEM | EPI_STORE_GLOBAL;
            op.n_chunks = (uint16_t)(k8 / 8);
            op.n_valid = (uint16_t)k;
            op.src_col = (uint16_t)du;
            op.row_out = (uint32_t)nd.row0;
            if (i == ns - 1) op.flags |= EPI_LAST_FWD_STORE;
            add_epi(op, 1, {{du, du + k8}}, {});              // ... and its stores: EPI stream 1
        }
        // ---- backward sweep (reverse post-order: every node after its ancestors)
        int ri = 0;
        for (int i = ns - 1; i >= 0; --i) {
            const TNode &nd = nodes[steps[i]];
            const int k = (int)nd.rows.size(), k8 = pad_to(k, 8);
            const int dx = dx0 + (ri % ring) * nmax;
            bool first = true;
            for (const Coupling &c : nd.coup) {
                const TNode &an = nodes[c.anc];
                const int m = (int)an.rows.size();
                const int k0 = c.r0 / 8 * 8, k1 = pad_to(c.r1, 8);
                MmaOp op{};
                op.d_col = (uint16_t)dx;
                op.a_hi_col = (uint16_t)(an.xs_off + k0);
                op.a_lo_col = (uint16_t)(xs_cols + an.xs_off + k0);
                op.n = (uint16_t)pad_to(k, 16);
                op.k8 = (uint16_t)((k1 - k0) / 8);
                op.flags = first ? 0 : MMA_ACCUMULATE;
                first = false;
                add_mma(op, [&](int nn, int kk) { return (nn < k && k0 + kk < m) ? -c.G[(size_t)(k0 + kk) * k + nn] : 0.0; });
            }
            EpiOp op{};
            op.flags = EPI_ADD_GLOBAL | EPI_STORE_GLOBAL | (i == ns - 1 ? EPI_AFTER_STORES : 0);
            op.n_chunks = (uint16_t)(k8 / 8);
            op.n_valid = (uint16_t)k;
            op.row_in = op.row_out = (uint32_t)nd.row0;
            std::vector<Range> rd, wr;
            int stream = 0;
            if (!first) {
                op.flags |= EPI_FROM_TMEM;
                op.src_col = (uint16_t)dx;
                rd.push_back({dx, dx + k8});
                stream = ri & 1;                                  // alternate with the result ring
                ++ri;
            }
            if (!nd.children.empty()) {
                op.flags |= EPI_ST_SPLIT;
                op.hi_col = (uint16_t)nd.xs_off;
                op.lo_col = (uint16_t)(xs_cols + nd.xs_off);
                wr.push_back({nd.xs_off, nd.xs_off + k8});
                wr.push_back({xs_cols + nd.xs_off, xs_cols + nd.xs_off + k8});
            }
            add_epi(op, stream, rd, wr);
        }
        if (!t.mma.empty()) t.mma.back().flags |= MMA_CHUNK_LAST;
        t.chunk_off.push_back((uint32_t)t.matrix.size());
        // the tile boundary is a full synchronisation: the last EPI op waits for the last MMA op, and the kernel
        // puts a barrier over all EPI warps behind the tile's ops (the MMA stream's first op of the next tile
        // waits for an EPI op of that tile)
        if (!t.mma.empty()) {
            const int last = (int)t.mma.size() - 1;
            if (epi_wait_mma_op.back() < last) {
                // only possible if the final node had no couplings; wait for everything anyway
                epi_wait_mma_op.back() = last;
                mma_commits[last] = 1;
            }
        }
        // ---- events
        std::vector<int> mma_evt(t.mma.size(), -1), epi_evt(t.epi.size(), -1);
        for (size_t m = 0; m < t.mma.size(); ++m) if (mma_commits[m]) mma_evt[m] = t.n_mma_events++;
        std::vector<int> epi_evt_read(t.epi.size(), -1);
        for (size_t e = 0; e < t.epi.size(); ++e) {
            if (epi_signals_read[e]) epi_evt_read[e] = t.n_epi_events++;
            if (epi_signals[e]) epi_evt[e] = t.n_epi_events++;
        }
        if (t.n_mma_events > TS_MAX_EVENTS || t.n_epi_events > TS_MAX_EVENTS) { err = "too many synchronisation events"; return false; }
        for (size_t m = 0; m < t.mma.size(); ++m) {
            t.mma[m].commit_mma = (int16_t)mma_evt[m];
            const std::array<int, 2> &w = mma_wait_epi_op[m];
            t.mma[m].wait_epi = (int16_t)(w[0] >= 0 ? (mma_wait_early[m][0] ? epi_evt_read : epi_evt)[w[0]] : -1);
            t.mma[m].wait_epi2 = (int16_t)(w[1] >= 0 ? (mma_wait_early[m][1] ? epi_evt_read : epi_evt)[w[1]] : -1);
        }
        for (size_t e = 0; e < t.epi.size(); ++e) {
            t.epi[e].signal_epi = (int16_t)epi_evt[e];
            t.epi[e].signal_read = (int16_t)epi_evt_read[e];
            t.epi[e].wait_mma = (int16_t)(epi_wait_mma_op[e] >= 0 ? mma_evt[epi_wait_mma_op[e]] : -1);
            t.epi[e].wait_epi = (int16_t)(epi_wait_epi_op[e] >= 0 ? (epi_wait_epi_early[e] ? epi_evt_read : epi_evt)[epi_wait_epi_op[e]] : -1);
        }
        return true;
    }

    bool run() {
        if (n > 2560) { err = "more than 2560 unknowns (dense block elimination not attempted)"; return false; }
        adj.assign(n, {});
        for (int c = 0; c < n; ++c)
            for (int q = p.m_colptr[c]; q < p.m_colptr[c + 1]; ++q) {
                const int r = p.m_rowidx[q];
                if (r != c) { adj[r].push_back(c); adj[c].push_back(r); }
            }
        std::vector<int> all(n);
        std::iota(all.begin(), all.end(), 0);
        const int root = bisect(all, -1);
        if (!err.empty()) return false;
        coarsen(root);
        std::vector<int> post;
        postorder(root, post);
        int row = 0;
        t.row_of_free.assign(n, -1);
        t.free_of_row.assign(n, -1);
        for (int v : post) {
            TNode &nd = nodes[v];
            if ((int)nd.rows.size() > TS_MAX_NODE) { err = "tree node of " + std::to_string(nd.rows.size()) + " rows"; return false; }
            nd.row0 = row;
            for (int f : nd.rows) { t.row_of_free[f] = row; t.free_of_row[row] = f; ++row; }
            for (int a = nd.parent; a >= 0; a = nodes[a].parent) nd.anc.push_back(a);
            for (int a : nd.anc) {
                nd.acc_off += pad_to((int)nodes[a].rows.size(), 16);
                nd.xs_off += pad_to((int)nodes[a].rows.size(), 8);
            }
            t.n_leaves += nd.children.empty();
        }
        t.n_nodes = (int)nodes.size();
        if (!factor(post)) return false;
        return emit(post);
    }
};

}  // namespace

void build_tensor_plan(HostPlan &p, int leaf_max) {
    leaf_max = std::max(8, std::min(leaf_max, TS_MAX_NODE));
    // Folding separators shortens the serial chain but deepens the accumulator stack; take the coarsest tree that
    // still leaves tensor memory for two source slots and a two-deep result ring (SDFA_TS_MERGE fixes the cap).
    std::vector<int> caps = {64, 56, 48, 40, 32, 0};
    if (const char *mc = std::getenv("SDFA_TS_MERGE")) caps = {std::atoi(mc)};
    std::string why;
    for (size_t i = 0; i < caps.size(); ++i) {
        p.tplan = TensorPlan();
        Builder b(p, leaf_max);
        b.merge_cap = caps[i];
        if (!b.run()) { why = b.err; continue; }
        if ((b.used_slots >= 2 && b.used_ring >= 2) || i + 1 == caps.size()) { p.tplan.valid = true; return; }
    }
    p.tplan = TensorPlan();
    p.tplan.why_not = why;
}

}  // namespace sdfa
