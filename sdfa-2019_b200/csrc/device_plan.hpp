// device_plan.hpp -- what sdfa_create uploads and what the kernel launchers in kernels.cu consume.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "plan.hpp"
#include <vector>

namespace sdfa {

// Where element (frame f, row, coordinate c) of the solve scratch lives:
//   (f / FL) * tile_stride + f % FL + row * row_stride + c * c_stride
// SIMT solve: FL = F frames per solve tile, rows of [3][F] floats ("tile-major"); tensor solve: FL = 128, one
// [n_free][128] matrix per (128-frame tile, coordinate).
struct ScratchLayout {
    int FL = 32;
    long long tile_stride = 0;
    int row_stride = 0, c_stride = 0;
};

struct DevicePlan {
    int device = -1;
    int n_verts = 0, n_tris = 0, n_cnsts = 0, n_free = 0, n_eq = 0;
    int sm_count = 0;
    // ---- assembly (K2)
    int n_asm_blocks = 0, asm_max_eq = 0, asm_max_rows = 0;
    int4           *asm_blocks = nullptr;   // {eq_begin, eq_end, row_begin, row_end}
    const float4   *asm_eq_meta = nullptr;  // per block-local equation: U0[3], U1[3], then the block rows of its corners (4 x int16)
    const int32_t  *asm_row_perm = nullptr;
    int4           *asm_walk = nullptr;     // per (block, warp) walk: {block-local equation or barrier / end mark, its source triangle,
                                            //  its slot group in the compact dgrad, -}
    const int32_t  *asm_warp_ptr = nullptr; // [blocks * 8 + 1]
    int asm_max_walk = 0;
    int out_fc = 8;                         // frames per CTA of k_output2 (8, 16 or 32): 8 keeps the set of lines written at any one time compact (1.41 -> 1.23 ms)
    int out_gen = 2;                        // output kernel: 2 = k_output2 (direct 16-byte loads), 1 = k_output (staged)
    int asm_gather_gen = 0;                 // assembly kernel for reference-layout dgrad: 3 = k_assemble_gather3 (TMA tensor-map boxes), 2 = k_assemble_gather2
                                            // (cp.async rings), 1 = k_assemble_gather, 0 = 3 for rows up to 1 MB (FLAME: 359 KB, 3.70 -> 2.82 ms), else 2 (the
                                            // twice-subdivided template's 5.7 MB rows measure 11.0 ms with boxes against 9.7 ms with the rings)
    int32_t        *asm_eq_src_local = nullptr; // source triangle per block-local equation
    const int32_t  *asm_row_ptr = nullptr;      // CSR incidence of the block rows (gather variant)
    const uint16_t *asm_inc = nullptr;
    int32_t        *eq_src = nullptr;       // equation block -> source triangle (>=0), -1 identity, -2 zero block
    int compact_stride = 0, compact_s_rows = 0;   // slots per frame of the compact dgrad; the rotation part starts at s_rows
    // ---- solve (K3)
    const uint8_t  *prog = nullptr;         // 16-byte aligned stage stream
    const uint32_t *stage_off = nullptr;
    const IoDesc   *io_desc = nullptr;
    const IoPhase  *io_phase = nullptr;
    int n_stages = 0, n_slots = 0, n_phases_fwd = 0, n_phases_bwd = 0, frames_per_tile = 32;
    // launch configuration decided once per handle (sdfa_create: configure_launches), never in function-local statics
    int solve_ctas_per_sm = 1;              // K3 (SIMT)
    int decode_max_clusters = 0;            // K1: resident CTA pairs
    long long *solve_prof = nullptr;        // optional cycle counters [sm_count*4][8] (SDFA_SOLVE_PROFILE=1)
    // ---- tensor-core solve (K3T, solve_tc.cu); used instead of K3 when use_tensor
    bool use_tensor = false;
    ScratchLayout layout;
    const MmaOp    *ts_mma = nullptr;
    const EpiOp    *ts_epi = nullptr;
    const uint8_t  *ts_matrix = nullptr;
    const uint32_t *ts_chunk_off = nullptr;
    int ts_n_mma = 0, ts_n_epi = 0, ts_n_chunks = 0, ts_n_mma_events = 0, ts_n_epi_events = 0, ts_n_ring_ops = 0;
    // ---- output (K5)
    // per 64-vertex chunk tables of the output kernel (rebuilt by upload_base when the base / constraints change):
    // out_full writes every vertex [n_verts,3] (the reference layout), out_free only the free vertices [n_free,3] in
    // ascending vertex order (sdfa_*_free entry points: the constrained rows are constants the caller already has)
    struct OutTables {
        int n_rows = 0;                     // vertices per frame this table set writes
        int16_t *line_of = nullptr;         // [chunks*192]
        float   *cval = nullptr;            // [chunks*192]
        int32_t *line_ptr = nullptr;        // [chunks+1]
        int32_t *line_off = nullptr;        // [n_free*3]
        float   *line_hi = nullptr, *line_lo = nullptr;
        int max_lines = 0;
    } out_full, out_free;
    // expansion free rows -> all vertices (sdfa_expand_free_dev): per output element the free-row element or -1 + constant
    int32_t *exp_src = nullptr;             // [n_verts*3]
    float   *exp_cval = nullptr;            // [n_verts*3]
    // ---- decode (K1)
    int k_scale = 0, k_rotat = 0;
    int n_pca_tris = 0;                     // source triangles of the basis given to sdfa_set_pca (rows / 6, rows / 3)
    float *wfull_scale = nullptr, *mfull_scale = nullptr, *wfull_rotat = nullptr, *mfull_rotat = nullptr;
    // tensor-core decode (decode_tc.cu): pre-split, pre-tiled basis images, bias and output offsets per row
    float *tc_w_scale = nullptr, *tc_w_rotat = nullptr;
    int tc_mt_scale = 0, tc_mt_rotat = 0;
    // second generation (decode_tc16.cu): scaled FP16 hi/lo basis images and 1 / s_w per part; used when decode_fp16 and the
    // basis widths fit its resident frames operands (tc16_fits), else the TF32 kernel
    const uint16_t *tc16_w_scale = nullptr, *tc16_w_rotat = nullptr;
    float tc16_inv_sw[2] = {1.f, 1.f};
    bool decode_fp16 = true, tc16_ready = false;
    int decode16_max_clusters = 0;
};

enum AssemblyMode { ASM_DGRAD = 0, ASM_MATRIX = 1 };

// All launchers are asynchronous on `stream` and return the cudaError_t of the launch.
// staged = dgrad is the block-planar compact buffer (decode output); else any [frame][tri][9] layout
cudaError_t launch_assembly(const DevicePlan &d, const float *dgrad, long long frame_stride, bool staged,
                            int n_frames, int mode, float *rhs, cudaStream_t stream);
cudaError_t launch_solve(const DevicePlan &d, float *scratch, int n_frames, cudaStream_t stream);
cudaError_t launch_solve_tc(const DevicePlan &d, float *scratch, int n_frames, cudaStream_t stream);
size_t solve_tc_smem_bytes(int n_mma, int n_epi);
size_t scratch_floats(const DevicePlan &d, int n_frames);
cudaError_t launch_output(const DevicePlan &d, const DevicePlan::OutTables &t, const float *scratch, int n_frames, float *out,
                          cudaStream_t stream);
cudaError_t launch_expand(const DevicePlan &d, const float *free_rows, int n_frames, float *out, cudaStream_t stream);
cudaError_t launch_decode_full(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                               float *dgrad_out, cudaStream_t stream);
// inverse.cu: verts_b holds n_frames meshes vb_stride floats apart; out is [n_frames, n_tris, 9] double or float
cudaError_t launch_deform_grad(const float *verts_a, const float *verts_b, long long vb_stride, const uint32_t *tris,
                               int n_tris, int n_frames, double eps, int as_matrix, void *out, bool out_f64,
                               cudaStream_t stream);
// inverse.cu: out[q] = (float)(a[q] * seq[lo[q]] + (1 - a[q]) * seq[hi[q]]) in float64, rows of `width` floats
cudaError_t launch_seek(const float *seq, long long width, const int2 *pairs, const double *weights, int n_query, float *out,
                        cudaStream_t stream);
// decode_tc.cu
cudaError_t launch_decode_tc(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                             float *ximg_scale, float *ximg_rotat, float *dgrad_out, cudaStream_t stream);
size_t tc_ximg_floats(int n_frames, int K);
// decode_tc16.cu
cudaError_t launch_decode_tc16(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                               float *ximg_scale, float *ximg_rotat, float *dgrad_out, cudaStream_t stream);
size_t tc16_ximg_floats(int n_frames, int K);
bool tc16_fits(int k_scale, int k_rotat);
int tc16_build_basis(const float *W, const float *mean, int K, const std::vector<int32_t> &rows_src, std::vector<uint16_t> &img,
                     float *inv_sw);
cudaError_t configure_decode_tc16(DevicePlan &d);
// Occupancy queries of the persistent kernels, stored in the plan (called by sdfa_create with the device current).
cudaError_t configure_solve(DevicePlan &d);
cudaError_t configure_decode_tc(DevicePlan &d);
size_t solve_smem_bytes(int n_slots, int frames_per_tile);
void count_launch();
long long launch_counter();

// rows_src[r] = row of the [*, K] basis W that GEMM row r (= compact slot r of the part) reproduces, or -1;
// returns the number of row tiles
int tc_build_basis(const float *W, const float *mean, int K, const std::vector<int32_t> &rows_src, std::vector<float> &img);
int tc_rows_per_tile();

}  // namespace sdfa
