// kernels.cu -- the per-frame CUDA path (sm_100a): assembly (K2), triangular solve (K3), constrained-vertex
// fill (K4) and a first decode kernel (K1).  Together they replace, per frame,
// TriangleDeformation::getMeshFromDeformationGradients (reference
// deformation/cpp/src/deform_triangle_impl.hpp:215-310) and, for K1, PcaInversion.forward +
// data_to_anime_feat (speech_anime/modules/output_module.py:115-116, speech_anime/model/model.py:246-257).
//
// Numerics: everything is float32 and works on the DISPLACEMENT from the identity deformation:
//   x = x_base + M^-1 A^T (T^T - I),  x_base = M^-1 A^T (stack(I) - A_r C) computed in fp64 on the host,
// which keeps the result within 1e-6 x bbox of the reference's fp64 path.
#include "device_plan.hpp"
#include "plan.hpp"

#include <atomic>
#include <cstdio>

namespace sdfa {

static std::atomic<long long> g_launches{0};
long long launch_counter() { return g_launches.load(); }

// =============================================================================================
// K2: per-equation transform + A^T (T^T - I) assembly.
//
// grid = (row blocks, frame lanes); a CTA owns one row block of one frame at a time.
//   phase 1  thread per block-local equation: E = R*S - I from the 9 dgrad values
//            (impl.hpp:226-244; rotation_log_exp::exp, rotation/utils_rotation.cpp:20-51), then the two
//            corner vectors g2 = E*U0, g3 = E*U1 and g1 = -(g2+g3) (coefficients of impl.hpp:106-116)
//            -> shared memory [eq][corner][3] (stride 9 words: conflict free)
//   phase 2  thread per row: sum the corner vectors incident to the row (CSR), write rhs[frame][row][3]
// E is evaluated without ever forming 1 + small:  E u = t + Q (u + t),  t = Es u,
//   Q v = a W v + b W (W v),  a = sin(th)/th,  b = (1 - cos th)/th^2 = 2 sin^2(th/2)/th^2.
struct AsmParams {
    const int4 *blocks;
    const int32_t *eq_id;
    const float *eq_u;
    const int32_t *row_perm, *row_ptr;
    const uint16_t *inc;
    const int32_t *eq_src;
    const float *dgrad;
    long long frame_stride;
    float *rhs;
    int n_frames, n_free, mode;
};

__device__ __forceinline__ void corner_vec(const float *d, float a, float b, const float *u, float *g) {
    // t = Es u (symmetric part, entries d0..d5 = s00,s01,s02,s11,s12,s22 minus identity)
    float t0 = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
    float t1 = d[1] * u[0] + d[3] * u[1] + d[4] * u[2];
    float t2 = d[2] * u[0] + d[4] * u[1] + d[5] * u[2];
    float s0 = u[0] + t0, s1 = u[1] + t1, s2 = u[2] + t2;
    // W = [[0,d6,d7],[-d6,0,d8],[-d7,-d8,0]]  (impl.hpp:232-235)
    float p0 = d[6] * s1 + d[7] * s2;
    float p1 = -d[6] * s0 + d[8] * s2;
    float p2 = -d[7] * s0 - d[8] * s1;
    float q0 = d[6] * p1 + d[7] * p2;
    float q1 = -d[6] * p0 + d[8] * p2;
    float q2 = -d[7] * p0 - d[8] * p1;
    g[0] = t0 + a * p0 + b * q0;
    g[1] = t1 + a * p1 + b * q1;
    g[2] = t2 + a * p2 + b * q2;
}

__global__ void __launch_bounds__(128) k_assemble(AsmParams P) {
    extern __shared__ float g_sh[];
    const int4 blk = P.blocks[blockIdx.x];
    const int n_eq = blk.y - blk.x, n_rows = blk.w - blk.z;
    for (int frame = blockIdx.y; frame < P.n_frames; frame += gridDim.y) {
        const float *row = P.dgrad + (long long)frame * P.frame_stride;
        for (int e = threadIdx.x; e < n_eq; e += blockDim.x) {
            const int ge = blk.x + e;
            const int src = P.eq_src[P.eq_id[ge]];
            float u0[3], u1[3], g2[3], g3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { u0[k] = __ldg(P.eq_u + ge * 6 + k); u1[k] = __ldg(P.eq_u + ge * 6 + 3 + k); }
            if (src >= 0) {
                float d[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) d[k] = __ldg(row + (long long)src * 9 + k);
                if (P.mode == ASM_DGRAD) {
                    float th2 = d[6] * d[6] + d[7] * d[7] + d[8] * d[8];
                    float th = sqrtf(th2);
                    float a = 0.f, b = 0.f;
                    if (th >= 1e-6f) {          // angle < 1e-6 => R = I (utils_rotation.cpp:46-47)
                        float sh = sinf(0.5f * th);
                        a = sinf(th) / th;
                        b = 2.f * sh * sh / th2;
                    }
                    corner_vec(d, a, b, u0, g2);
                    corner_vec(d, a, b, u1, g3);
                } else {                        // raw row-major T (impl.hpp:391-397): E = T - I
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        g2[c] = (d[3 * c] - (c == 0 ? 1.f : 0.f)) * u0[0] + (d[3 * c + 1] - (c == 1 ? 1.f : 0.f)) * u0[1] +
                                (d[3 * c + 2] - (c == 2 ? 1.f : 0.f)) * u0[2];
                        g3[c] = (d[3 * c] - (c == 0 ? 1.f : 0.f)) * u1[0] + (d[3 * c + 1] - (c == 1 ? 1.f : 0.f)) * u1[1] +
                                (d[3 * c + 2] - (c == 2 ? 1.f : 0.f)) * u1[2];
                    }
                }
            } else if (src == -1) {             // identity block (impl.hpp:264-268): T - I = 0
#pragma unroll
                for (int c = 0; c < 3; ++c) g2[c] = g3[c] = 0.f;
            } else {                            // block left at zero by setZero (impl.hpp:224): T = 0, E = -I
#pragma unroll
                for (int c = 0; c < 3; ++c) { g2[c] = -u0[c]; g3[c] = -u1[c]; }
            }
            float *g = g_sh + e * 9;
#pragma unroll
            for (int c = 0; c < 3; ++c) { g[c] = -(g2[c] + g3[c]); g[3 + c] = g2[c]; g[6 + c] = g3[c]; }
        }
        __syncthreads();
        for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
            const int gr = blk.z + r;
            const int q0 = P.row_ptr[gr], q1 = P.row_ptr[gr + 1];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int q = q0; q < q1; ++q) {
                const float *g = g_sh + 3 * (int)P.inc[q];
                s0 += g[0]; s1 += g[1]; s2 += g[2];
            }
            float *dst = P.rhs + ((long long)frame * P.n_free + P.row_perm[gr]) * 3;
            dst[0] = s0; dst[1] = s1; dst[2] = s2;
        }
        __syncthreads();
    }
}

cudaError_t launch_assembly(const DevicePlan &d, const float *dgrad, long long frame_stride, const int32_t *eq_src,
                            int n_frames, int mode, float *rhs, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    AsmParams P{d.asm_blocks, d.asm_eq_id, d.asm_eq_u, d.asm_row_perm, d.asm_row_ptr, d.asm_inc, eq_src,
                dgrad, frame_stride, rhs, n_frames, d.n_free, mode};
    size_t smem = (size_t)d.asm_max_eq * 9 * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)d.n_asm_blocks, (unsigned)(n_frames < 32768 ? n_frames : 32768));
    k_assemble<<<grid, 128, smem, stream>>>(P);
    g_launches++;
    return cudaGetLastError();
}

// =============================================================================================
// K3: batched multi-RHS sparse triangular solve (forward + backward) -- interpreter of the solve program
// built by schedule.cpp.  One CTA = one tile of 32 frames (lane = frame, 3 coordinates per lane).
//   warp 0            producer: streams the program's stages global -> shared ring with cp.async.bulk
//                     (TMA), completion on mbarriers
//   warps 1..NCW      consumers: interpret the ops; rows of a level are dealt round-robin to the warps,
//                     levels are separated by a named barrier over the consumer warps only
// The state (piece + root-path rows, 396 B per row) lives in shared memory; finished rows are spilled to
// the rhs scratch in global memory (L2 resident) and re-read by the backward sweep.
constexpr int RING = 4;
constexpr int NCW = 8;                          // consumer warps
constexpr int SOLVE_THREADS = 32 * (NCW + 1);

struct SolveParams {
    const uint8_t *prog;
    const uint32_t *stage_off;
    int n_stages, n_slots;
    float *rhs;                                 // [n_frames][n_free][3] in: rhs, scratch: y
    float *out;                                 // [n_frames][n_verts][3]
    const int32_t *row_vert;
    const float *xb_hi, *xb_lo;
    int n_frames, n_free, n_verts, n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }

__device__ __forceinline__ void run_row_task(const uint8_t *task, uint8_t *state_lane) {
    const uint4 th = *reinterpret_cast<const uint4 *>(task);       // TaskHeader
    const int n = (int)(th.y & 0xFFFFFFu);
    const uint4 *e = reinterpret_cast<const uint4 *>(task + 16);   // two entries per uint4
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll 4
    for (int k = 0; k < n; k += 2) {
        const uint4 p = e[k >> 1];
        const float c0 = __uint_as_float(p.x), c1 = __uint_as_float(p.z);
        const float *s0 = reinterpret_cast<const float *>(state_lane + p.y);
        const float *s1 = reinterpret_cast<const float *>(state_lane + p.w);
        a0 = fmaf(c0, s0[0], a0); a1 = fmaf(c0, s0[COORD_STRIDE], a1); a2 = fmaf(c0, s0[2 * COORD_STRIDE], a2);
        b0 = fmaf(c1, s1[0], b0); b1 = fmaf(c1, s1[COORD_STRIDE], b1); b2 = fmaf(c1, s1[2 * COORD_STRIDE], b2);
    }
    float *t = reinterpret_cast<float *>(state_lane + th.x);
    const float dinv = __uint_as_float(th.z);                      // 1.0 for partial (non-final) rows
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (!(th.y & TASK_OVERWRITE)) { v0 = t[0]; v1 = t[COORD_STRIDE]; v2 = t[2 * COORD_STRIDE]; }
    t[0] = (v0 - (a0 + b0)) * dinv;
    t[COORD_STRIDE] = (v1 - (a1 + b1)) * dinv;
    t[2 * COORD_STRIDE] = (v2 - (a2 + b2)) * dinv;
}

__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(SolveParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RING * STAGE_BYTES);   // full[RING], empty[RING]
    float *state = reinterpret_cast<float *>(smem + RING * STAGE_BYTES + 2 * RING * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // zero the state: padding entries of row tasks multiply slot 0 by 0.0, which must not be NaN
    for (int i = threadIdx.x; i < P.n_slots * SLOT_WORDS; i += blockDim.x) state[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);            // full: the producer's arrive.expect_tx
            mbar_init(smem_u32(&bars[RING + s]), NCW);   // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int s = 0; s < P.n_stages; ++s, ++it) {
                    const uint32_t slot = it % RING, phase = (it / RING) & 1u;
                    mbar_wait(smem_u32(&bars[RING + slot]), phase ^ 1u);
                    const uint32_t off = P.stage_off[s], bytes = P.stage_off[s + 1] - off;
                    const uint32_t full = smem_u32(&bars[slot]);
                    mbar_arrive_expect_tx(full, bytes);
                    tma_bulk_g2s(smem_u32(ring + slot * STAGE_BYTES), P.prog + off, bytes, full);
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    const int cw = warp - 1;
    uint8_t *state_lane = reinterpret_cast<uint8_t *>(state + lane);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int frame0 = tile * FRAMES_PER_TILE;
        const int nvalid = min(FRAMES_PER_TILE, P.n_frames - frame0);
        for (int s = 0; s < P.n_stages; ++s, ++it) {
            const uint32_t slot = it % RING, phase = (it / RING) & 1u;
            mbar_wait(smem_u32(&bars[slot]), phase);
            const uint8_t *stage = ring + slot * STAGE_BYTES;
            const int n_ops = (int)reinterpret_cast<const uint32_t *>(stage)[0];
            uint32_t at = 16;
            for (int o = 0; o < n_ops; ++o) {
                const uint4 hw = *reinterpret_cast<const uint4 *>(stage + at);   // OpHeader
                const uint32_t type = hw.x & 0xFFFFu, flags = hw.x >> 16;
                if (flags & OPF_SYNC_BEFORE) consumer_bar();
                if (type == OP_ROWS) {
                    const uint32_t *table = reinterpret_cast<const uint32_t *>(stage + hw.z);
                    for (uint32_t t = cw; t < hw.y; t += NCW) run_row_task(stage + table[t], state_lane);
                    at = hw.w;
                } else {
                    const int row0 = (int)hw.y, n_rows = (int)hw.z, ne = 3 * n_rows;
                    const uint32_t *table = reinterpret_cast<const uint32_t *>(stage + hw.w);
                    for (int f = cw; f < nvalid; f += NCW) {
                        if (type == OP_LOAD) {
                            const float *src = P.rhs + ((long long)(frame0 + f) * P.n_free + row0) * 3;
                            for (int e = lane; e < ne; e += 32) {
                                const int r = e / 3, c = e - 3 * r;
                                const uint32_t w = table[r];
                                float *dst = state + (w & 0xFFFFFFu) + c * COORD_STRIDE + f;
                                float v = src[e];
                                if (w & LOAD_ADD_BIT) v += *dst;
                                *dst = v;
                            }
                        } else if (type == OP_STORE_Y) {
                            float *dst = P.rhs + ((long long)(frame0 + f) * P.n_free + row0) * 3;
                            for (int e = lane; e < ne; e += 32) {
                                const int r = e / 3, c = e - 3 * r;
                                dst[e] = state[(table[r] & 0xFFFFFFu) + c * COORD_STRIDE + f];
                            }
                        } else {   // OP_STORE_X
                            float *dst = P.out + (long long)(frame0 + f) * P.n_verts * 3;
                            for (int e = lane; e < ne; e += 32) {
                                const int r = e / 3, c = e - 3 * r;
                                const float v = state[(table[r] & 0xFFFFFFu) + c * COORD_STRIDE + f];
                                const int row = row0 + r;
                                dst[(long long)__ldg(P.row_vert + row) * 3 + c] =
                                    __ldg(P.xb_hi + row * 3 + c) + (__ldg(P.xb_lo + row * 3 + c) + v);
                            }
                        }
                    }
                    at = hw.w + (((uint32_t)n_rows * 4u + 15u) & ~15u);
                }
                if (flags & OPF_SYNC_AFTER) consumer_bar();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[RING + slot]));
        }
    }
}

size_t solve_smem_bytes(int n_slots) {
    return (size_t)RING * STAGE_BYTES + 2 * RING * 8 + (size_t)n_slots * SLOT_BYTES;
}

cudaError_t launch_solve(const DevicePlan &d, float *rhs_scratch, int n_frames, float *out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const size_t smem = solve_smem_bytes(d.n_slots);
    static int configured_device = -1;
    static size_t configured_smem = 0;
    static int ctas_per_sm = 1;
    if (configured_device != d.device || configured_smem != smem) {
        cudaError_t e = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_solve, SOLVE_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
        configured_device = d.device;
        configured_smem = smem;
    }
    const int n_tiles = (n_frames + FRAMES_PER_TILE - 1) / FRAMES_PER_TILE;
    SolveParams P{d.prog, d.stage_off, d.n_stages, d.n_slots, rhs_scratch, out, d.row_vert, d.xbase_hi, d.xbase_lo,
                  n_frames, d.n_free, d.n_verts, n_tiles};
    int grid = d.sm_count * ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    k_solve<<<grid, SOLVE_THREADS, smem, stream>>>(P);
    g_launches++;
    return cudaGetLastError();
}

// =============================================================================================
// K4: constrained vertices are copied through unchanged (impl.hpp:302-308).
__global__ void k_fill_constraints(const int32_t *cnst_vert, const float *cnst_pos, int n_cnsts, int n_verts,
                                   int n_frames, float *out) {
    const int ne = n_cnsts * 3;
    for (int frame = blockIdx.y; frame < n_frames; frame += gridDim.y) {
        float *dst = out + (long long)frame * n_verts * 3;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
            const int i = e / 3, c = e - 3 * i;
            dst[(long long)__ldg(cnst_vert + i) * 3 + c] = __ldg(cnst_pos + e);
        }
    }
}

cudaError_t launch_fill_constraints(const DevicePlan &d, int n_frames, float *out, cudaStream_t stream) {
    if (n_frames <= 0 || d.n_cnsts == 0) return cudaSuccess;
    int bx = (d.n_cnsts * 3 + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)(n_frames < 32768 ? n_frames : 32768));
    k_fill_constraints<<<grid, 256, 0, stream>>>(d.cnst_vert, d.cnst_pos, d.n_cnsts, d.n_verts, n_frames, out);
    g_launches++;
    return cudaGetLastError();
}

// =============================================================================================
// K1 (first version, exact fp32 FMA on CUDA cores): dgrad[f][tri][0..5] = coeff_s[f] . Ws[tri*6+s] + ms,
// dgrad[f][tri][6..8] = coeff_r[f] . Wr[tri*3+r] + mr  -- F.linear x2 + the scale/rotation interleave.
// 64 frames x 64 outputs per CTA, 4x4 per thread, K streamed through shared memory in chunks of 16.
constexpr int DEC_TF = 64, DEC_TJ = 64, DEC_TK = 16;

__global__ void __launch_bounds__(256) k_decode(const float *__restrict__ coeff, int K, const float *__restrict__ W,
                                                const float *__restrict__ mean, int J, int per_tri, int col0,
                                                int n_frames, float *__restrict__ out, long long out_stride) {
    __shared__ float xs[DEC_TK][DEC_TF + 4];
    __shared__ float ws[DEC_TK][DEC_TJ + 4];
    const int f0 = blockIdx.y * DEC_TF, j0 = blockIdx.x * DEC_TJ;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // tx: outputs, ty: frames
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += DEC_TK) {
        for (int i = threadIdx.x; i < DEC_TK * DEC_TF; i += 256) {
            const int kk = i % DEC_TK, r = i / DEC_TK;           // consecutive threads walk k: coalesced rows
            const int k = k0 + kk;
            xs[kk][r] = (k < K && f0 + r < n_frames) ? coeff[(long long)(f0 + r) * K + k] : 0.f;
            ws[kk][r] = (k < K && j0 + r < J) ? W[(long long)(j0 + r) * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DEC_TK; ++kk) {
            float xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { xv[i] = xs[kk][ty * 4 + i]; wv[i] = ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int f = f0 + ty * 4 + i;
        if (f >= n_frames) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj >= J) continue;
            const int tri = jj / per_tri, s = jj - tri * per_tri;
            out[(long long)f * out_stride + (long long)tri * 9 + col0 + s] = acc[i][j] + mean[jj];
        }
    }
}

cudaError_t launch_decode(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                          bool full_layout, float *dgrad_out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int ntri = full_layout ? d.n_tris : d.n_needed;
    const float *ws = full_layout ? d.wfull_scale : d.w_scale, *ms = full_layout ? d.mfull_scale : d.m_scale;
    const float *wr = full_layout ? d.wfull_rotat : d.w_rotat, *mr = full_layout ? d.mfull_rotat : d.m_rotat;
    const long long stride = (long long)ntri * 9;
    dim3 gs((unsigned)((ntri * 6 + DEC_TJ - 1) / DEC_TJ), (unsigned)((n_frames + DEC_TF - 1) / DEC_TF));
    k_decode<<<gs, 256, 0, stream>>>(coeff_scale, d.k_scale, ws, ms, ntri * 6, 6, 0, n_frames, dgrad_out, stride);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 gr((unsigned)((ntri * 3 + DEC_TJ - 1) / DEC_TJ), (unsigned)((n_frames + DEC_TF - 1) / DEC_TF));
    k_decode<<<gr, 256, 0, stream>>>(coeff_rotat, d.k_rotat, wr, mr, ntri * 3, 3, 6, n_frames, dgrad_out, stride);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace sdfa
