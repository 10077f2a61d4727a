// analysis.cpp -- one-time host analysis of a template (fp64): equation coefficients, the SPD system
// matrix M = A^T A + reg*I, a fill-reducing ordering, the sparse Cholesky factor and base solutions.
//
// This is the B200 build's counterpart of TriangleDeformation::setStaticTarget
// (reference deformation/cpp/src/deform_triangle_impl.hpp:7-142).  The reference factors the same
// matrix with Eigen::SparseLU (deform_triangle.hpp:27); M is symmetric positive definite (reg > 0 or
// >= 1 constrained vertex per component), so a Cholesky factor L L^T = P M P^T solves the same system
// and its two triangular sweeps are what the GPU solve kernel executes per frame.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <set>

namespace sdfa {

namespace {

// Classical Gram-Schmidt QR of the 3x2 edge matrix with the reference's degenerate guard
// (_qrFactorize, impl.hpp:479-511: |v| < 1e-6 -> r = 1, q = 0), then U = R^-1 Q^T (impl.hpp:100).
void triangle_frame(const float *v1, const float *v2, const float *v3, double *u /*[6]*/) {
    constexpr double EPS = 1e-6;
    double e1[3], e2[3];
    for (int d = 0; d < 3; ++d) {
        // the reference subtracts Eigen::Vector3f's, i.e. in float32, and then widens (impl.hpp:92-97)
        float a = v2[d] - v1[d], b = v3[d] - v1[d];
        e1[d] = (double)a;
        e2[d] = (double)b;
    }
    double q0[3], q1[3];
    double r00 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    if (r00 < EPS) { r00 = 1.0; q0[0] = q0[1] = q0[2] = 0.0; }
    else { for (int d = 0; d < 3; ++d) q0[d] = e1[d] / r00; }
    double r01 = q0[0] * e2[0] + q0[1] * e2[1] + q0[2] * e2[2];
    double v[3] = {e2[0] - r01 * q0[0], e2[1] - r01 * q0[1], e2[2] - r01 * q0[2]};
    double r11 = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (r11 < EPS) { r11 = 1.0; q1[0] = q1[1] = q1[2] = 0.0; }
    else { for (int d = 0; d < 3; ++d) q1[d] = v[d] / r11; }
    // R = [[r00, r01], [0, r11]]  =>  R^-1 = [[1/r00, -r01/(r00 r11)], [0, 1/r11]]
    for (int d = 0; d < 3; ++d) {
        u[d]     = q0[d] / r00 - (r01 / (r00 * r11)) * q1[d];
        u[3 + d] = q1[d] / r11;
    }
}

inline void corner_coef(const double *u, int corner, double *c /*[3]*/) {
    // rows 3k+r of A: v1 gets -U(0,r)-U(1,r), v2 gets U(0,r), v3 gets U(1,r)  (impl.hpp:106-116)
    for (int r = 0; r < 3; ++r)
        c[r] = corner == 0 ? (-u[r] - u[3 + r]) : (corner == 1 ? u[r] : u[3 + r]);
}

struct Trip { int r, c; double v; };

}  // namespace

int build_system(HostPlan &p, std::string &err) {
    const int nv = p.n_verts, nt = p.n_tris, nc = p.n_cnsts;
    if (nv <= 0 || nt <= 0 || nc < 0 || nc > nv) { err = "bad counts"; return 1; }
    for (int i = 0; i < nt * 3; ++i)
        if (p.tris[i] >= (uint32_t)nv) { err = "triangle index out of range"; return 1; }
    p.vi_to_free.assign(nv, -1);
    p.vi_to_cnst.assign(nv, -1);
    for (int i = 0; i < nc; ++i) {
        uint32_t ci = p.cnsts[i];
        if (ci >= (uint32_t)nv) { err = "constraint index out of range"; return 1; }
        if (p.vi_to_cnst[ci] >= 0) { err = "duplicate constraint index (reference asserts, impl.hpp:60)"; return 1; }
        p.vi_to_cnst[ci] = i;                        // constrained vertex c[i] -> column i of A_r
    }
    p.free_to_vi.clear();
    for (int v = 0; v < nv; ++v)
        if (p.vi_to_cnst[v] < 0) { p.vi_to_free[v] = (int)p.free_to_vi.size(); p.free_to_vi.push_back(v); }
    p.n_free = (int)p.free_to_vi.size();
    if (p.n_free == 0) { err = "all vertices are constrained: nothing to solve"; return 1; }

    // equation blocks (impl.hpp:18-22, 102-103)
    p.eq_tri.clear();
    if (!p.corr_count.empty() && (int)p.corr_count.size() != nt) { err = "corr_count must have n_tris entries"; return 1; }
    for (int j = 0; j < nt; ++j) {
        int steps = p.corr_count.empty() ? 1 : std::max(1, (int)p.corr_count[j]);
        for (int s = 0; s < steps; ++s) p.eq_tri.push_back(j);
    }
    p.n_eq = (int)p.eq_tri.size();

    p.tri_u.assign((size_t)nt * 6, 0.0);
    for (int j = 0; j < nt; ++j) {
        const uint32_t *t = &p.tris[(size_t)j * 3];
        triangle_frame(&p.verts[(size_t)t[0] * 3], &p.verts[(size_t)t[1] * 3], &p.verts[(size_t)t[2] * 3],
                       &p.tri_u[(size_t)j * 6]);
    }

    // active equation blocks and M = A^T A (lower triangle) from the 3x3 corner Gram matrix of each block
    p.active_eq.clear();
    std::vector<Trip> trips;
    trips.reserve((size_t)p.n_eq * 6);
    for (int k = 0; k < p.n_eq; ++k) {
        int j = p.eq_tri[k];
        const uint32_t *t = &p.tris[(size_t)j * 3];
        int fr[3] = {p.vi_to_free[t[0]], p.vi_to_free[t[1]], p.vi_to_free[t[2]]};
        if (fr[0] < 0 && fr[1] < 0 && fr[2] < 0) continue;
        p.active_eq.push_back(k);
        double c[3][3];
        for (int q = 0; q < 3; ++q) corner_coef(&p.tri_u[(size_t)j * 6], q, c[q]);
        for (int a = 0; a < 3; ++a) {
            if (fr[a] < 0) continue;
            for (int b = 0; b < 3; ++b) {
                if (fr[b] < 0) continue;
                if (fr[a] < fr[b]) continue;         // keep lower triangle (row >= col)
                double d = c[a][0] * c[b][0] + c[a][1] * c[b][1] + c[a][2] * c[b][2];
                trips.push_back({fr[a], fr[b], d});
            }
        }
    }
    p.n_active = (int)p.active_eq.size();
    for (int i = 0; i < p.n_free; ++i) trips.push_back({i, i, p.reg});   // impl.hpp:126-131
    std::sort(trips.begin(), trips.end(), [](const Trip &a, const Trip &b) {
        return a.c != b.c ? a.c < b.c : a.r < b.r;
    });
    p.m_colptr.assign(p.n_free + 1, 0);
    p.m_rowidx.clear();
    p.m_val.clear();
    for (size_t i = 0; i < trips.size();) {
        size_t j = i;
        double s = 0.0;
        while (j < trips.size() && trips[j].c == trips[i].c && trips[j].r == trips[i].r) s += trips[j++].v;
        p.m_rowidx.push_back(trips[i].r);
        p.m_val.push_back(s);
        p.m_colptr[trips[i].c + 1]++;
        i = j;
    }
    for (int c = 0; c < p.n_free; ++c) p.m_colptr[c + 1] += p.m_colptr[c];
    return 0;
}

// ------------------------------------------------------------------------------------------
// Minimum-degree ordering on the explicit elimination graph (exact external degree, ties by index).
// n is at most ~1e5 and the graphs are planar-ish meshes, so the O(sum deg^2) cost is negligible
// next to everything else in sdfa_create.
static std::vector<int> minimum_degree(int n, const std::vector<int> &colptr, const std::vector<int> &rowidx) {
    std::vector<std::vector<int>> adj(n);
    for (int c = 0; c < n; ++c)
        for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
            int r = rowidx[q];
            if (r != c) { adj[r].push_back(c); adj[c].push_back(r); }
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::set<std::pair<int, int>> heap;
    for (int v = 0; v < n; ++v) heap.insert({(int)adj[v].size(), v});
    std::vector<char> done(n, 0);
    std::vector<int> order, tmp;
    order.reserve(n);
    while (!heap.empty()) {
        int v = heap.begin()->second;
        heap.erase(heap.begin());
        done[v] = 1;
        order.push_back(v);
        std::vector<int> nb;
        nb.swap(adj[v]);
        for (int u : nb) {
            heap.erase({(int)adj[u].size(), u});
            tmp.clear();
            tmp.reserve(adj[u].size() + nb.size());
            std::set_union(adj[u].begin(), adj[u].end(), nb.begin(), nb.end(), std::back_inserter(tmp));
            adj[u].clear();
            for (int w : tmp) if (w != u && w != v) adj[u].push_back(w);
            heap.insert({(int)adj[u].size(), u});
        }
    }
    return order;
}

// upper-triangular CSC of P M P^T: column k holds rows i <= k
static void permuted_upper(const HostPlan &p, const std::vector<int> &iperm, std::vector<int> &cp,
                           std::vector<int> &ri, std::vector<double> &vv) {
    int n = p.n_free;
    cp.assign(n + 1, 0);
    for (int c = 0; c < n; ++c)
        for (int q = p.m_colptr[c]; q < p.m_colptr[c + 1]; ++q) cp[std::max(iperm[p.m_rowidx[q]], iperm[c]) + 1]++;
    for (int c = 0; c < n; ++c) cp[c + 1] += cp[c];
    ri.assign(cp[n], 0);
    vv.assign(cp[n], 0.0);
    std::vector<int> fill(cp.begin(), cp.end() - 1);
    for (int c = 0; c < n; ++c)
        for (int q = p.m_colptr[c]; q < p.m_colptr[c + 1]; ++q) {
            int i = iperm[p.m_rowidx[q]], j = iperm[c];
            int col = std::max(i, j), row = std::min(i, j);
            ri[fill[col]] = row;
            vv[fill[col]++] = p.m_val[q];
        }
}

static std::vector<int> elimination_tree(int n, const std::vector<int> &cp, const std::vector<int> &ri) {
    std::vector<int> parent(n, -1), anc(n, -1);
    for (int k = 0; k < n; ++k)
        for (int q = cp[k]; q < cp[k + 1]; ++q) {
            int i = ri[q];
            while (i != -1 && i < k) {
                int nxt = anc[i];
                anc[i] = k;
                if (nxt == -1) parent[i] = k;
                i = nxt;
            }
        }
    return parent;
}

int order_and_factor(HostPlan &p, std::string &err) {
    const int n = p.n_free;
    std::vector<int> order = minimum_degree(n, p.m_colptr, p.m_rowidx);   // order[new] = old
    std::vector<int> iperm(n);
    for (int i = 0; i < n; ++i) iperm[order[i]] = i;
    std::vector<int> cp, ri;
    std::vector<double> vv;
    permuted_upper(p, iperm, cp, ri, vv);
    std::vector<int> parent = elimination_tree(n, cp, ri);
    // postorder so that every subtree is a contiguous index range (children in increasing order)
    {
        std::vector<int> head(n, -1), next(n, -1), post;
        post.reserve(n);
        for (int j = n - 1; j >= 0; --j)
            if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
        std::vector<int> stack;
        for (int r = 0; r < n; ++r) {
            if (parent[r] >= 0) continue;
            stack.push_back(r);
            while (!stack.empty()) {
                int v = stack.back();
                int c = head[v];
                if (c >= 0) { head[v] = next[c]; stack.push_back(c); }
                else { post.push_back(v); stack.pop_back(); }
            }
        }
        std::vector<int> order2(n);
        for (int i = 0; i < n; ++i) order2[i] = order[post[i]];
        order.swap(order2);
        for (int i = 0; i < n; ++i) iperm[order[i]] = i;
        permuted_upper(p, iperm, cp, ri, vv);
        parent = elimination_tree(n, cp, ri);
    }
    p.perm = order;
    p.iperm = iperm;
    p.parent = parent;

    // symbolic: pattern of row k of L = nodes reached from the entries of column k of the upper part by
    // walking up the elimination tree (stop at already-marked nodes)
    std::vector<int> mark(n, -1), count(n, 1), pat;
    auto reach = [&](int k) {
        pat.clear();
        mark[k] = k;
        for (int q = cp[k]; q < cp[k + 1]; ++q) {
            int i = ri[q];
            while (i < k && mark[i] != k) { pat.push_back(i); mark[i] = k; i = parent[i]; }
        }
        std::sort(pat.begin(), pat.end());
    };
    for (int k = 0; k < n; ++k) { reach(k); for (int j : pat) count[j]++; }
    p.l_colptr.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) p.l_colptr[j + 1] = p.l_colptr[j] + count[j];
    p.l_rowidx.assign(p.l_colptr[n], 0);
    p.l_val.assign(p.l_colptr[n], 0.0);
    std::vector<int> fill(n);
    for (int j = 0; j < n; ++j) fill[j] = p.l_colptr[j] + 1;        // slot 0 of a column = diagonal
    std::fill(mark.begin(), mark.end(), -1);
    std::vector<double> x(n, 0.0);
    // numeric up-looking Cholesky: row k of L solves L[0:k,0:k] l = M[0:k,k]
    for (int k = 0; k < n; ++k) {
        reach(k);
        double d = 0.0;
        for (int q = cp[k]; q < cp[k + 1]; ++q) {
            if (ri[q] == k) d += vv[q];
            else x[ri[q]] += vv[q];
        }
        for (int j : pat) {
            double lkj = x[j] / p.l_val[p.l_colptr[j]];
            x[j] = 0.0;
            for (int q = p.l_colptr[j] + 1; q < fill[j]; ++q) x[p.l_rowidx[q]] -= p.l_val[q] * lkj;
            d -= lkj * lkj;
            p.l_rowidx[fill[j]] = k;
            p.l_val[fill[j]++] = lkj;
        }
        if (!(d > 0.0) || !std::isfinite(d)) {
            err = "A^T A + reg*I is not positive definite at pivot " + std::to_string(k);
            return 2;
        }
        p.l_rowidx[p.l_colptr[k]] = k;
        p.l_val[p.l_colptr[k]] = std::sqrt(d);
    }
    return 0;
}

void solve_factored(const HostPlan &p, std::vector<double> &b) {
    const int n = p.n_free;
    for (int j = 0; j < n; ++j) {                       // L y = b
        double dj = p.l_val[p.l_colptr[j]];
        for (int c = 0; c < 3; ++c) b[(size_t)j * 3 + c] /= dj;
        for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) {
            int i = p.l_rowidx[q];
            double l = p.l_val[q];
            for (int c = 0; c < 3; ++c) b[(size_t)i * 3 + c] -= l * b[(size_t)j * 3 + c];
        }
    }
    for (int j = n - 1; j >= 0; --j) {                  // L^T x = y
        double s[3] = {b[(size_t)j * 3], b[(size_t)j * 3 + 1], b[(size_t)j * 3 + 2]};
        for (int q = p.l_colptr[j] + 1; q < p.l_colptr[j + 1]; ++q) {
            int i = p.l_rowidx[q];
            double l = p.l_val[q];
            for (int c = 0; c < 3; ++c) s[c] -= l * b[(size_t)i * 3 + c];
        }
        double dj = p.l_val[p.l_colptr[j]];
        for (int c = 0; c < 3; ++c) b[(size_t)j * 3 + c] = s[c] / dj;
    }
}

// x_base = M^-1 A^T (stack(I) - A_r C): the solution for the identity deformation with constraint
// positions C.  The GPU adds the per-frame displacement M^-1 A^T (T^T - I) to it, which
// keeps the fp32 kernels inside the 1e-6 x bbox tolerance.
void compute_base_solution(HostPlan &p, const float *cnst_pos) {
    const int n = p.n_free;
    p.cnst_pos.assign((size_t)p.n_cnsts * 3, 0.f);
    for (int i = 0; i < p.n_cnsts; ++i)
        for (int c = 0; c < 3; ++c)
            p.cnst_pos[(size_t)i * 3 + c] = cnst_pos ? cnst_pos[(size_t)i * 3 + c] : p.verts[(size_t)p.cnsts[i] * 3 + c];
    std::vector<double> rhs((size_t)n * 3, 0.0);
    for (int k : p.active_eq) {
        int j = p.eq_tri[k];
        const uint32_t *t = &p.tris[(size_t)j * 3];
        double c[3][3];
        for (int q = 0; q < 3; ++q) corner_coef(&p.tri_u[(size_t)j * 6], q, c[q]);
        // b_k = I - sum_{constrained corner q} coef_q (outer) C_q   (3x3, rows r)
        double b[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int q = 0; q < 3; ++q) {
            int ci = p.vi_to_cnst[t[q]];
            if (ci < 0) continue;
            for (int r = 0; r < 3; ++r)
                for (int d = 0; d < 3; ++d) b[r][d] -= c[q][r] * (double)p.cnst_pos[(size_t)ci * 3 + d];
        }
        for (int a = 0; a < 3; ++a) {
            int fi = p.vi_to_free[t[a]];
            if (fi < 0) continue;
            int row = p.iperm[fi];
            for (int d = 0; d < 3; ++d)
                rhs[(size_t)row * 3 + d] += c[a][0] * b[0][d] + c[a][1] * b[1][d] + c[a][2] * b[2][d];
        }
    }
    solve_factored(p, rhs);
    p.x_base.swap(rhs);
}

}  // namespace sdfa
