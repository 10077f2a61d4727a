// tplan.hpp -- the tensor-core solve plan (kernel K3T, solve_tc.cu).
//
// The reference solves (A^T A + reg I) X = A^T B with Eigen::SparseLU per frame
// (deformation/cpp/src/deform_triangle_impl.hpp:286-292).  Here, for templates whose system fits, the two
// triangular sweeps are restated as a short sequence of small dense products on the 5th-generation tensor
// cores, with the right-hand sides on the MMA's M dimension (128 columns of the batch = 128 TMEM lanes) and
// the system rows on N / K:
//
//   nested dissection tree of M's graph (geometric bisection of the free vertices); for a tree node J
//   (a leaf patch or a separator) with the coupled rows B(J) of its ancestors, F = Schur-updated blocks:
//       P_J = F_JJ^-1,      G_J = F_BJ P_J
//   forward  (post-order):   t_B -= G_J t_J,   u_J = P_J t_J          (t_J = rhs rows + updates)
//   backward (pre-order):    x_J = u_J - G_J^T x_B
//
// which is block Cholesky with the diagonal solves folded into explicit inverses of the (<= 64 x 64) diagonal
// blocks -- every step is D[128 cols, N] (+)= A[128 cols, K] * Bt[N, K] with A and D in tensor memory and the
// fixed matrices Bt streamed from L2 as pre-split TF32 hi/lo tile images (3xTF32: A_hi B_hi + A_lo B_hi +
// A_hi B_lo, fp32 accumulation).
//
// The plan is three instruction streams per 128-column tile plus the matrix byte stream:
//   MMA stream (one issuing thread)               : MmaOp  -- a product group on tensor-memory column ranges
//   two EPI streams (8 warps each, thread = column): EpiOp  -- global rows <-> tensor memory, TF32 hi/lo splitting
// synchronised by single-use-per-tile mbarrier events (an op names at most one event per other stream to wait
// for and at most one to signal; the planner derives them from the column hazards).  The forward sweep's loads
// (rhs rows -> operand slots) run on EPI stream 0 and its stores (u rows) on stream 1; the backward sweep's
// x ops alternate between the streams with the result ring, so that the tensor-memory round trips of
// independent tree nodes overlap.  (One EPI stream -- SDFA_TS_STREAMS=1 -- is the round-1 arrangement.)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sdfa {

constexpr int TS_COLS = 128;                  // right-hand-side columns per tile (= TMEM lanes = UMMA M)
constexpr int TS_TMEM_COLS = 512;
constexpr int TS_MAX_NODE = 64;               // rows per tree node (K and N of a product)
constexpr int TS_STAGE_BYTES = 32768;         // matrix ring stage: hi + lo images of <= 64 x 64
constexpr int TS_MAX_EVENTS = 256;            // per direction

enum EpiFlags : uint16_t {
    EPI_FROM_TMEM    = 1,     // v  = tmem[src_col + i]
    EPI_ADD_GLOBAL   = 2,     // v += scratch[row_in + i]
    EPI_STORE_GLOBAL = 4,     // scratch[row_out + i] = v
    EPI_ST_RAW       = 8,     // tmem[hi_col + i] = v
    EPI_ST_SPLIT     = 16,    // tmem[hi_col + i] = tf32(v), tmem[lo_col + i] = tf32(v - tf32(v))
    EPI_LAST_FWD_STORE = 32,  // the forward sweep's last store: after it the rows it wrote may be bulk-loaded
    EPI_AFTER_STORES = 64,    // the first op that adds rows stored earlier in the tile (start of the backward sweep)
    EPI_ZERO_SRC     = 128,   // tmem[src_col + i] = 0 after reading (accumulator columns handed to the next subtree)
};

struct EpiOp {                // 32 bytes
    int16_t  wait_mma;        // MMA event to wait for before touching tensor memory (-1: none)
    int16_t  signal_epi;      // EPI event to signal when done (-1: none)
    uint16_t n_chunks;        // 8-column chunks
    uint16_t n_valid;         // rows that exist (the rest of the chunks is zero padding)
    uint16_t src_col, hi_col, lo_col, flags;
    uint32_t row_in, row_out; // scratch rows (tensor order)
    uint16_t stream;          // EPI stream (0 / 1) that executes the op
    int16_t  wait_epi;        // event of the OTHER EPI stream to wait for (-1: none)
    uint16_t ring_seq;        // ops that add or store rows: position among them in the tile (row-ring stage = running count % stages)
    int16_t  signal_read;     // EPI event to signal as soon as the op has READ its tensor-memory source (-1: none): the
                              // columns may be overwritten while the op still stores rows / writes its own results
};

enum MmaFlags : uint16_t {
    MMA_ACCUMULATE  = 1,      // first product accumulates onto D (else overwrites)
    MMA_CHUNK_FIRST = 2,      // first op of a matrix chunk: wait until the chunk has landed
    MMA_CHUNK_LAST  = 4,      // last op of a matrix chunk: release the ring stage afterwards
};

struct MmaOp {                // 32 bytes
    int16_t  wait_epi;        // EPI event to wait for before issuing (-1: none)
    int16_t  commit_mma;      // MMA event committed after this op (-1: none)
    uint16_t d_col, a_hi_col, a_lo_col;
    uint16_t n;               // N (multiple of 16)
    uint16_t k8;              // K / 8
    uint16_t flags;
    uint32_t b_hi_off, b_lo_off;   // byte offsets of the hi / lo tile images inside the ring stage (1024-aligned)
    int16_t  wait_epi2;       // a second EPI event (of the other EPI stream) to wait for (-1: none)
    uint16_t pad;
    uint32_t reserved;
};

struct TensorPlan {
    bool valid = false;
    std::string why_not;                  // reason when !valid
    std::vector<int> row_of_free;         // free column -> scratch row (nested-dissection order)
    std::vector<int> free_of_row;
    std::vector<MmaOp> mma;
    std::vector<EpiOp> epi;
    std::vector<uint8_t> matrix;          // chunk payloads back to back (each 1024-aligned)
    std::vector<uint32_t> chunk_off;      // n_chunks + 1
    int n_mma_events = 0, n_epi_events = 0;
    int n_streams = 1;                    // EPI streams the ops are dealt to
    int n_ring_ops = 0;                   // ops per tile that add or store scratch rows (each owns a row-ring stage)
    int n_nodes = 0, n_leaves = 0;
    long long nk_products = 0;            // sum of N*K over all product groups (x3 MMAs each)
    int tmem_fwd = 0, tmem_bwd = 0;       // columns used by the two sweeps
};

struct HostPlan;
// Builds p.tplan; leaves valid = false (with why_not) when the template does not fit the scheme
// (too many unknowns, pinned pivots of an unconstrained template, a separator path wider than tensor memory).
void build_tensor_plan(HostPlan &p, int leaf_max);

// [rows x 32] K-major SWIZZLE_128B tile image indexing (float index), shared by host packers and kernels
inline int ts_swz(int r, int k) { return (r >> 3) * 256 + (r & 7) * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3)); }

}  // namespace sdfa
