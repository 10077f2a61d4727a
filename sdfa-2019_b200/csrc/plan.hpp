// plan.hpp -- host-side analysis products and the binary formats the kernels consume.
//
// Everything here is computed once per template by sdfa_create() (the B200-native replacement of
// TriangleDeformation::setStaticTarget, reference deformation/cpp/src/deform_triangle_impl.hpp:7-142)
// in fp64 on the host and then uploaded; the per-frame work is CUDA only (kernels.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tplan.hpp"

namespace sdfa {

// ------------------------------------------------------------------------------------------
// Geometry of the solve tile.  One CTA of the solve kernel owns a tile of F frames (lane = frame),
// F = 32 when the resident rows fit in shared memory, else 16 or 8 (large factors: config 5).
// A state "slot" holds one row of the permuted system for those frames: [3 coords][F] floats, which is
// also the layout of one row in the global scratch ("tile-major": scratch[tile][row][coord][frame]), so
// rows move between global and shared memory with plain TMA bulk copies.
constexpr int MAX_FRAMES_PER_TILE = 32;
constexpr int slot_words(int f) { return 3 * f; }
constexpr int slot_bytes(int f) { return 12 * f; }

// ------------------------------------------------------------------------------------------
// Solve program.  Two sweeps (forward L y = b, backward L^T x = y), each a sequence of PHASES; a phase
// handles one "piece" (a contiguous range of supernodes of the postordered elimination tree):
//     data loads (TMA, IO warp)  ->  levels of row tasks (consumer warps)  ->  data stores (TMA, IO warp)
// The task lists are a byte stream cut into stages of at most STAGE_BYTES that the streamer warp copies
// global -> shared with cp.async.bulk into a ring; the IO warp follows the separate IoPhase/IoDesc tables.
constexpr int STAGE_BYTES = 8192;

enum OpType : uint16_t {
    OP_ROWS        = 1,   // a = n_tasks, b = byte offset (in stage) of the u32 task-offset table,
                          // c = byte offset of the next op
    OP_PHASE_BEGIN = 2,   // wait until the phase's rows have landed in shared memory
    OP_PHASE_END   = 3,   // make the phase's results visible to the async proxy and signal the IO warp
};
enum OpFlags : uint16_t {
    OPF_SYNC_AFTER = 2,    // consumer barrier after the op (end of a level)
};
struct OpHeader {        // 16 bytes, 16-byte aligned inside the stage
    uint16_t type, flags;
    uint32_t a, b, c;
};
struct StageHeader {     // first 16 bytes of every stage
    uint32_t n_ops, bytes, reserved0, reserved1;
};

// Tasks of OP_ROWS: a 16-byte header followed by the entries.
//   kind A (one row):   header {target, -, -, n | flags}; n (even) entries of 8 bytes {coeff, src}
//        acc = sum coeff * slot[src];   slot[target] = (OVERWRITE ? 0 : slot[target]) - acc
//   kind B (TASK_GROUP, up to 3 rows reading the same sources): header {target0, target1, target2, n | flags},
//        n entries of 16 bytes {src, c0, c1, c2}; every source value is loaded once and used for all rows.
//        Absent rows have target 0xFFFFFFFF; OVERWRITE of row r is flag bit 25 + r.
struct TaskHeader {
    uint32_t target_byte_off, target1, target2;
    uint32_t n_entries_flags;   // low 24 bits count, bit 24 = TASK_GROUP, bits 25.. = OVERWRITE per row
};
constexpr uint32_t TASK_GROUP = 1u << 24, TASK_OVERWRITE = 1u << 25;
constexpr int TASK_BATCH_B = 4;      // kind-B entry lists are padded to a multiple of this many sources
struct TaskEntry { float coeff; uint32_t src_byte_off; };
struct GroupEntry { uint32_t src_byte_off; float c[3]; };

// IO tables: n_rows consecutive permuted rows <-> n_rows consecutive slots, one cp.async.bulk each.
struct IoDesc { uint32_t row, n_rows, slot, reserved; };
struct IoPhase { uint32_t load_begin, load_end, store_begin, store_end; };   // ranges in the IoDesc array

struct SolveProgram {
    std::vector<uint8_t>  bytes;        // all stages back to back, each padded to a multiple of 16
    std::vector<uint32_t> stage_off;    // byte offset of each stage, n_stages + 1 entries
    std::vector<IoDesc>   io_desc;
    std::vector<IoPhase>  io_phase;     // forward phases then backward phases
    int n_phases_fwd = 0, n_phases_bwd = 0;
    int frames_per_tile = 32;           // F
    int n_slots = 0;                    // state slots a CTA needs (peak over both sweeps)
    int n_steps_fwd = 0, n_steps_bwd = 0, n_supernodes = 0;
    long long n_entries = 0;            // useful multiply-adds per (frame, coordinate)
    long long n_entries_padded = 0;     // issued ones (zero padding of grouped rows included)
};

// ------------------------------------------------------------------------------------------
// Assembly plan (kernel K2): the free rows are grouped into row blocks; a CTA handles one (row block, tile of 32
// frames) with lane = frame: a warp evaluates an equation's two corner vectors for the 32 frames and adds them
// to the block's row accumulators in shared memory.  The block's equations are coloured so that one colour never
// touches a row twice, and the colours are processed in order with a barrier in between: no atomics, a fixed
// summation order.
constexpr int COMPACT_TILE = 64;                 // frames per tile of the compact dgrad: K2 handles two frames per lane (packed fp32x2)
constexpr int ASM_MAX_COLOURS = 32;
#ifndef ASM_WARPS_N
#define ASM_WARPS_N 16
#endif
constexpr int ASM_WARPS_PER_BLOCK = ASM_WARPS_N;
constexpr int16_t ASM_SCHED_BARRIER = -1, ASM_SCHED_END = -2;
struct AssemblyBlock {
    int eq_begin, eq_end;       // range in eq_* arrays (block-local equations, duplicates across blocks allowed)
    int row_begin, row_end;     // range in row_perm
    int n_colours;
};
struct AssemblyPlan {
    std::vector<AssemblyBlock> blocks;
    // per block-local equation (sorted by colour): which equation block it is, the frame of its target triangle
    // U0[3], U1[3] (rows of U = R^-1 Q^T, impl.hpp:98-100) and the block-local rows of its three corners
    std::vector<int32_t> eq_id;         // global equation-block index (0..n_eq)
    std::vector<float>   eq_u;          // 8 floats per entry (6 used)
    std::vector<int16_t> eq_rows;       // 4 per entry: block row of corner 0, 1, 2 (-1: not a row of this block), pad
    std::vector<int32_t> colour_ptr;    // ASM_MAX_COLOURS + 1 per block: first block-local equation of each colour
    std::vector<int32_t> row_perm;      // per block-local row: scratch row (where to write)
    std::vector<int32_t> row_ptr;       // rows + 1: incidence ranges (gather variant: per-row sums, frame at a time)
    std::vector<uint16_t> inc;          // incidence: block-local equation * 3 + corner
    std::vector<int16_t> warp_sched;    // per (block, warp): block-local equations, ASM_SCHED_BARRIER between colours, ASM_SCHED_END
    std::vector<int32_t> warp_ptr;      // [blocks * ASM_WARPS_PER_BLOCK + 1] start of each walk in warp_sched
    int max_eq_per_block = 0, max_rows_per_block = 0;
    // Frame-tiled compact dgrad (what the decode kernel writes and the staged assembly reads):
    // [tile of COMPACT_TILE frames][slot][COMPACT_TILE frames]; active equation u (slot_eq[u]) owns the scale slots
    // 6 u .. 6 u + 5 (s00,s01,s02,s11,s12,s22) and the rotation slots compact_s_rows + 3 u .. + 2 (r01,r02,r12);
    // both parts are padded to whole GEMM row tiles, so a decode GEMM row is simply its slot.
    std::vector<int32_t> eq_slot;       // per block-local equation: its u
    std::vector<int32_t> slot_eq;       // u -> equation block (the active equations in block-walk order)
    int compact_s_rows = 0;
    int compact_stride = 0;             // slots = floats per frame
};

// ------------------------------------------------------------------------------------------
struct HostPlan {
    // template
    int n_verts = 0, n_tris = 0, n_cnsts = 0, n_free = 0, n_eq = 0, n_active = 0;
    double reg = 1e-10;
    std::vector<float>    verts;        // [n_verts*3]
    std::vector<uint32_t> tris;         // [n_tris*3]
    std::vector<uint32_t> cnsts;        // [n_cnsts]
    std::vector<uint32_t> corr_count;   // [] or [n_tris]
    // reference column maps (impl.hpp:36-73)
    std::vector<int> vi_to_free, vi_to_cnst, free_to_vi;
    std::vector<int> eq_tri;            // equation block -> target triangle
    std::vector<double> tri_u;          // per triangle: U0[3], U1[3] in fp64
    std::vector<int> active_eq;         // equation blocks touching >= 1 free vertex
    // system matrix M = A^T A + reg I (lower triangle, CSC, original free-column order)
    std::vector<int> m_colptr, m_rowidx;
    std::vector<double> m_val;
    // ordering and factor: perm[new] = old free column; L in CSC (diagonal first in every column)
    std::vector<int> perm, iperm, parent;
    std::vector<int> l_colptr, l_rowidx;
    std::vector<double> l_val;
    // base solution for the current constraint positions: M^-1 A^T (stack(I) - A_r C), permuted order
    std::vector<double> x_base;         // [n_free*3]
    std::vector<float>  cnst_pos;       // [n_cnsts*3] currently active constraint positions
    SolveProgram prog;
    AssemblyPlan asmplan;
    // tensor-core solve (tplan.hpp); when use_tensor is set the scratch rows follow its nested-dissection
    // order instead of the Cholesky permutation
    TensorPlan tplan;
    bool use_tensor = false;
    std::vector<int> scratch_row;       // free column -> row of the solve scratch (= iperm, or tplan.row_of_free)
};

// analysis.cpp
int  build_system(HostPlan &p, std::string &err);                 // maps, U, M
int  order_and_factor(HostPlan &p, std::string &err);             // perm, etree, L
void compute_base_solution(HostPlan &p, const float *cnst_pos);   // x_base
void solve_factored(const HostPlan &p, std::vector<double> &rhs_perm /* [n_free*3] in/out */);
// schedule.cpp
void build_solve_program(HostPlan &p, int piece_cap, int supernode_cap, int subtree_cap, int frames_per_tile);
void build_assembly_plan(HostPlan &p, int rows_per_block, int max_eq_per_block);
constexpr int ASM_ROWS_MAX = 116; // vertices per row block of the assembly kernel (api.cpp splits evenly below this)
constexpr int ASM_MAX_EQ = 512;   // equations per row block the assembly kernel keeps in registers/shared memory

}  // namespace sdfa
