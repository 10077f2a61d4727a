// kernels.cu -- the per-frame CUDA path (sm_100a): assembly (K2), triangular solve (K3), constrained-vertex
// fill (K4) and a first decode kernel (K1).  Together they replace, per frame,
// TriangleDeformation::getMeshFromDeformationGradients (reference
// deformation/cpp/src/deform_triangle_impl.hpp:215-310) and, for K1, PcaInversion.forward +
// data_to_anime_feat (speech_anime/modules/output_module.py:115-116, speech_anime/model/model.py:246-257).
//
// Numerics: everything is float32 and works on the DISPLACEMENT from the identity deformation:
//   x = x_base + M^-1 A^T (T^T - I),  x_base = M^-1 A^T (stack(I) - A_r C) computed in fp64 on the host,
// which keeps the result within 1e-6 x bbox of the reference's fp64 path.
#include <cuda.h>

#include "device_plan.hpp"
#include "plan.hpp"

#include <algorithm>
#include <atomic>
#include <cstdio>

namespace sdfa {

static std::atomic<long long> g_launches{0};
long long launch_counter() { return g_launches.load(); }
void count_launch() { g_launches++; }

// ---- mbarrier / TMA bulk-copy helpers (shared by K2 and K3) -------------------------------------------
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
__device__ __forceinline__ void tma_bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// =============================================================================================
// K2: per-equation transform + A^T (T^T - I) assembly, lane = frame.
//
// grid = (row blocks, tiles of 32 frames); a CTA owns one row block for one tile.  A warp takes an equation of
// the block and evaluates, for its 32 frames at once (lane = frame),
//      E = R*S - I from the 9 dgrad values (impl.hpp:226-244; rotation_log_exp::exp, rotation/utils_rotation.cpp:20-51)
//      the corner vectors g2 = E*U0, g3 = E*U1, g1 = -(g2+g3)   (coefficients of impl.hpp:106-116)
// and adds them to the rows of the block's accumulator acc[row][xyz][frame] in shared memory (conflict free: a
// lane only ever touches its own frame).  The host coloured the block's equations so that one colour never
// touches a row twice (schedule.cpp); colours run in order with a block barrier in between, so there are no
// atomics and the summation order is fixed.  At the end the accumulator rows leave as the solve scratch's
// rows, one 128-byte line per warp instruction.
// E is evaluated without ever forming 1 + small:  E u = t + Q (u + t),  t = Es u,
//   Q v = a W v + b W (W v),  a = sin(th)/th,  b = (1 - cos th)/th^2 = 2 sin^2(th/2)/th^2.
// Input: the decode kernel's frame-tiled compact buffer [tile][slot][32] (nine coalesced lines per equation);
// k_assemble_gather below takes any [frame][triangle][9] tensor (the reference's dgrad layout) instead.
struct AsmParams {
    const int4 *blocks;                      // {eq_begin, eq_end, row_begin, row_end}
    const int4 *walk;                        // per (block, warp): {equation | ASM_SCHED_BARRIER | ASM_SCHED_END, source triangle, slot group, -}
    const int32_t *warp_ptr;
    const float4 *eq_meta;                   // 2 per block-local equation: U0, U1, corner rows
    const int32_t *row_perm;
    const int32_t *eq_src_local;             // source triangle per block-local equation (gather variant)
    const int32_t *row_ptr;                  // CSR incidence of the block rows (gather variant)
    const uint16_t *inc;
    int max_eq;
    const float *dgrad;
    long long frame_stride;                  // gather variant: floats per frame; else slots per frame
    int s_rows;                              // first rotation slot of the compact buffer
    float *rhs;
    int n_frames, mode, max_rows, max_walk;
    ScratchLayout L;
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

// The two corner vectors g2 = E*U0, g3 = E*U1 of one equation for one frame, E = R*S - I from the nine dgrad values
// (mode ASM_DGRAD) or E = T - I from a raw row-major matrix (ASM_MATRIX).
__device__ __forceinline__ void eq_vectors(int mode, const float (&d)[9], const float (&u0)[3], const float (&u1)[3],
                                           float (&g2)[3], float (&g3)[3]) {
    if (mode == ASM_DGRAD) {
        const float th2 = d[6] * d[6] + d[7] * d[7] + d[8] * d[8];
        float a = 0.f, b = 0.f;
        if (th2 >= 1e-12f) {        // angle < 1e-6 => R = I (utils_rotation.cpp:46-47)
            if (th2 <= 1.f) {
                // Taylor series in th^2 (remainder < 3e-8 for th <= 1): no sqrt, sin or division
                a = 1.f - th2 * (1.f / 6.f) * (1.f - th2 * (1.f / 20.f) * (1.f - th2 * (1.f / 42.f) * (1.f - th2 * (1.f / 72.f))));
                b = 0.5f - th2 * (1.f / 24.f) * (1.f - th2 * (1.f / 30.f) * (1.f - th2 * (1.f / 56.f) * (1.f - th2 * (1.f / 90.f))));
            } else {
                const float th = sqrtf(th2);
                const float sh2 = sinf(0.5f * th);
                a = sinf(th) / th;
                b = 2.f * sh2 * sh2 / th2;
            }
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
}

// One equation for one frame (this lane's): its corner vectors added to the accumulator rows of its corners.
// m0, m1 = the equation record: U0, U1 and the block rows of the three corners.
__device__ __forceinline__ void eq_apply(float *acc, int lane, int mode, int src, const float (&d)[9], float4 m0, float4 m1) {
    if (src == -1) return;                                       // identity block (impl.hpp:264-268): T - I = 0
    const float u0[3] = {m0.x, m0.y, m0.z}, u1[3] = {m0.w, m1.x, m1.y};
    const uint32_t r01 = __float_as_uint(m1.z), r23 = __float_as_uint(m1.w);
    const int rx = (short)(r01 & 0xFFFFu), ry = (short)(r01 >> 16), rz = (short)(r23 & 0xFFFFu);
    float g2[3], g3[3];
    if (src >= 0) eq_vectors(mode, d, u0, u1, g2, g3);
    else {                              // block left at zero by setZero (impl.hpp:224): T = 0, E = -I
#pragma unroll
        for (int c = 0; c < 3; ++c) { g2[c] = -u0[c]; g3[c] = -u1[c]; }
    }
    // corner 0 (v1) gets -(g2+g3), corner 1 (v2) g2, corner 2 (v3) g3
    if (rx >= 0) {
        float *t = acc + rx * 96 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * 32] -= g2[c] + g3[c];
    }
    if (ry >= 0) {
        float *t = acc + ry * 96 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * 32] += g2[c];
    }
    if (rz >= 0) {
        float *t = acc + rz * 96 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * 32] += g3[c];
    }
}

constexpr int ASM_WARPS = ASM_WARPS_PER_BLOCK;
constexpr int ASM_THREADS = 32 * ASM_WARPS;

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 of sm_100): a lane carries two frames
__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// corner_vec for two frames at once; nd6..nd8 = -d6..-d8
__device__ __forceinline__ void corner_vec2(const float2 *d, float2 nd6, float2 nd7, float2 nd8, float2 a, float2 b, const float *u, float2 *g) {
    const float2 u0 = f2(u[0]), u1 = f2(u[1]), u2 = f2(u[2]);
    const float2 t0 = fma2(d[2], u2, fma2(d[1], u1, mul2(d[0], u0)));
    const float2 t1 = fma2(d[4], u2, fma2(d[3], u1, mul2(d[1], u0)));
    const float2 t2 = fma2(d[5], u2, fma2(d[4], u1, mul2(d[2], u0)));
    const float2 s0 = add2(u0, t0), s1 = add2(u1, t1), s2 = add2(u2, t2);
    const float2 p0 = fma2(d[7], s2, mul2(d[6], s1));
    const float2 p1 = fma2(d[8], s2, mul2(nd6, s0));
    const float2 p2 = fma2(nd8, s1, mul2(nd7, s0));
    const float2 q0 = fma2(d[7], p2, mul2(d[6], p1));
    const float2 q1 = fma2(d[8], p2, mul2(nd6, p0));
    const float2 q2 = fma2(nd8, p1, mul2(nd7, p0));
    g[0] = fma2(b, q0, fma2(a, p0, t0));
    g[1] = fma2(b, q1, fma2(a, p1, t1));
    g[2] = fma2(b, q2, fma2(a, p2, t2));
}

// a = sin(th)/th, b = (1 - cos th)/th^2 for one frame (th2 = th^2); see eq_vectors
__device__ __forceinline__ void rot_coeffs(float th2, float &a, float &b) {
    a = b = 0.f;
    if (th2 < 1e-12f) return;           // angle < 1e-6 => R = I (utils_rotation.cpp:46-47)
    const float th = sqrtf(th2), sh2 = sinf(0.5f * th);
    a = sinf(th) / th;
    b = 2.f * sh2 * sh2 / th2;
}

// One equation for the two frames of this lane (k_assemble): corner vectors added to the accumulator rows of its corners;
// the same arithmetic as eq_apply, packed.
__device__ __forceinline__ void eq_apply2(float *acc, int lane, int mode, int src, const float2 (&d)[9], float4 m0, float4 m1) {
    if (src == -1) return;                                       // identity block (impl.hpp:264-268): T - I = 0
    const float u0[3] = {m0.x, m0.y, m0.z}, u1[3] = {m0.w, m1.x, m1.y};
    const uint32_t r01 = __float_as_uint(m1.z), r23 = __float_as_uint(m1.w);
    const int rx = (short)(r01 & 0xFFFFu), ry = (short)(r01 >> 16), rz = (short)(r23 & 0xFFFFu);
    float2 g2[3], g3[3];
    if (src >= 0) {
        if (mode == ASM_DGRAD) {
            const float2 th2 = fma2(d[8], d[8], fma2(d[7], d[7], mul2(d[6], d[6])));
            // Taylor series in th^2 (remainder < 3e-8 for th <= 1): no sqrt, sin or division
            const float2 one = f2(1.f);
            float2 a = fma2(mul2(th2, f2(-1.f / 72.f)), one, one);
            a = fma2(mul2(th2, f2(-1.f / 42.f)), a, one);
            a = fma2(mul2(th2, f2(-1.f / 20.f)), a, one);
            a = fma2(mul2(th2, f2(-1.f / 6.f)), a, one);
            float2 b = fma2(mul2(th2, f2(-1.f / 90.f)), one, one);
            b = fma2(mul2(th2, f2(-1.f / 56.f)), b, one);
            b = fma2(mul2(th2, f2(-1.f / 30.f)), b, one);
            b = fma2(mul2(th2, f2(-1.f / 24.f)), b, f2(0.5f));
            // out of the series' range (angle < 1e-6 => R = I, or th > 1): per frame, rarely
            if (th2.x < 1e-12f || th2.x > 1.f) rot_coeffs(th2.x, a.x, b.x);
            if (th2.y < 1e-12f || th2.y > 1.f) rot_coeffs(th2.y, a.y, b.y);
            const float2 nd6 = neg2(d[6]), nd7 = neg2(d[7]), nd8 = neg2(d[8]);
            corner_vec2(d, nd6, nd7, nd8, a, b, u0, g2);
            corner_vec2(d, nd6, nd7, nd8, a, b, u1, g3);
        } else {                        // raw row-major T (impl.hpp:391-397): E = T - I
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 e0 = add2(d[3 * c], f2(c == 0 ? -1.f : 0.f)), e1 = add2(d[3 * c + 1], f2(c == 1 ? -1.f : 0.f)),
                             e2 = add2(d[3 * c + 2], f2(c == 2 ? -1.f : 0.f));
                g2[c] = fma2(e2, f2(u0[2]), fma2(e1, f2(u0[1]), mul2(e0, f2(u0[0]))));
                g3[c] = fma2(e2, f2(u1[2]), fma2(e1, f2(u1[1]), mul2(e0, f2(u1[0]))));
            }
        }
    } else {                            // block left at zero by setZero (impl.hpp:224): T = 0, E = -I
#pragma unroll
        for (int c = 0; c < 3; ++c) { g2[c] = f2(-u0[c]); g3[c] = f2(-u1[c]); }
    }
    // corner 0 (v1) gets -(g2+g3), corner 1 (v2) g2, corner 2 (v3) g3
    if (rx >= 0) {
        float2 *t = reinterpret_cast<float2 *>(acc + rx * 3 * COMPACT_TILE) + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * (COMPACT_TILE / 2)] = add2(t[c * (COMPACT_TILE / 2)], neg2(add2(g2[c], g3[c])));
    }
    if (ry >= 0) {
        float2 *t = reinterpret_cast<float2 *>(acc + ry * 3 * COMPACT_TILE) + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * (COMPACT_TILE / 2)] = add2(t[c * (COMPACT_TILE / 2)], g2[c]);
    }
    if (rz >= 0) {
        float2 *t = reinterpret_cast<float2 *>(acc + rz * 3 * COMPACT_TILE) + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c * (COMPACT_TILE / 2)] = add2(t[c * (COMPACT_TILE / 2)], g3[c]);
    }
}

static_assert(COMPACT_TILE == 64, "k_assemble carries two frames per lane");

#ifndef ASM_L2_AHEAD
#define ASM_L2_AHEAD 4
#endif
#ifndef ASM_MIN_BLOCKS
#define ASM_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(ASM_THREADS, ASM_MIN_BLOCKS) k_assemble(AsmParams P) {
    extern __shared__ __align__(16) float acc[];                     // [row][3][64]
    constexpr int CT = COMPACT_TILE;
    const int4 blk = P.blocks[blockIdx.x];
    const int n_rows = blk.w - blk.z;
    const int tile = blockIdx.y;
    const int frame0 = tile * CT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_rows * 3 * CT; i += ASM_THREADS) acc[i] = 0.f;
    // this tile's lines (six scale lines and three rotation lines per equation); the lane's two frames are adjacent
    const float2 *in = reinterpret_cast<const float2 *>(P.dgrad + (long long)tile * P.frame_stride * CT) + lane;
    // the block's walks (one per warp) go to shared memory first, so that an entry costs a shared-memory read and the only
    // long-latency loads are an equation's values and its record
    int4 *walk_sh = reinterpret_cast<int4 *>(acc + P.max_rows * 3 * CT);
    {
        const int w0 = P.warp_ptr[blockIdx.x * ASM_WARPS], w1 = P.warp_ptr[blockIdx.x * ASM_WARPS + ASM_WARPS];
        for (int i = threadIdx.x; i < w1 - w0; i += ASM_THREADS) walk_sh[i] = P.walk[w0 + i];
    }
    const int4 *walk = walk_sh + (P.warp_ptr[blockIdx.x * ASM_WARPS + warp] - P.warp_ptr[blockIdx.x * ASM_WARPS]);
    struct Eq { int e, src; float4 m0, m1; float2 d[9]; };
    auto fetch = [&](int4 ent, Eq &q) {
        q.e = ent.x; q.src = ent.y;
        if (ent.x < 0) return;
        q.m0 = __ldg(P.eq_meta + (size_t)(blk.x + ent.x) * 2);
        q.m1 = __ldg(P.eq_meta + (size_t)(blk.x + ent.x) * 2 + 1);
        // slots of identity / zero blocks hold zeros: always readable
        const float2 *qs = in + (size_t)ent.z * 6 * (CT / 2), *qr = in + ((size_t)P.s_rows + (size_t)ent.z * 3) * (CT / 2);
#pragma unroll
        for (int j = 0; j < 6; ++j) q.d[j] = __ldcs(qs + j * (CT / 2));
#pragma unroll
        for (int j = 0; j < 3; ++j) q.d[6 + j] = __ldcs(qr + j * (CT / 2));
    };
    // the equation's corner vectors for this lane's two frames, added to the block rows of its three corners
    auto apply = [&](const Eq &q) { eq_apply2(acc, lane, P.mode, q.src, q.d, q.m0, q.m1); };
    // The warp walks its equations one at a time: fetch (nine 256-byte lines), apply.  What hides the memory latency is
    // (a) 16 warps per CTA, two CTAs per SM -- a second equation in registers (software pipelining) costs 30 registers and
    // with them a third of the warps: 2.08 ms against 1.66 ms -- and (b) an L2 prefetch of the equation ASM_L2_AHEAD
    // entries further down the walk (18 lines: one prefetch instruction, lane = line), which costs no registers.
    Eq q;
    __syncthreads();                                                 // accumulator zeroed, walks in shared memory
    const int4 *pf = walk;
    auto l2_prefetch = [&]() {
        const int4 ent = *pf;
        if (ent.x == ASM_SCHED_END) return;
        ++pf;
        if (ent.x < 0 || lane >= 18) return;
        const float *line = lane < 12 ? P.dgrad + ((long long)tile * P.frame_stride + (long long)ent.z * 6) * CT + lane * 32
                                      : P.dgrad + ((long long)tile * P.frame_stride + P.s_rows + (long long)ent.z * 3) * CT + (lane - 12) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    };
#pragma unroll 1
    for (int i = 0; i < ASM_L2_AHEAD; ++i) l2_prefetch();
    for (;;) {
        fetch(*walk++, q);
        l2_prefetch();
        if (q.e == ASM_SCHED_END) break;
        if (q.e == ASM_SCHED_BARRIER) __syncthreads(); else apply(q);
    }
    // write-out: the lane's two frames are adjacent in a scratch line as well (frames per line is even)
    const int fr = frame0 + 2 * lane;
    float *dst_tile = P.rhs + (long long)(fr / P.L.FL) * P.L.tile_stride + fr % P.L.FL;
    const float2 *acc2 = reinterpret_cast<const float2 *>(acc) + lane;
    for (int line = warp; line < n_rows * 3; line += ASM_WARPS) {
        const int r = line / 3, c = line - 3 * r;
        float2 v = acc2[line * (CT / 2)];
        if (fr >= P.n_frames) v.x = 0.f;
        if (fr + 1 >= P.n_frames) v.y = 0.f;
        *reinterpret_cast<float2 *>(dst_tile + (long long)P.row_perm[blk.z + r] * P.L.row_stride + c * P.L.c_stride) = v;
    }
}

// Gather variant for any [frame][triangle][9] tensor (the reference's dgrad layout, possibly with correspondences).
// The values of one frame sit in one 359 KB row, 36 bytes per active triangle, so this variant walks the tile frame
// by frame with the whole CTA on one row at a time (neighbouring DRAM pages): thread = equation, its nine values
// gathered with 4-byte cp.async copies two frames ahead into a double-buffered stage, corner vectors into shared
// memory, then thread = row sums the incident corner vectors (CSR, fixed order, no atomics) into a
// [row*3+c][33] transpose buffer; one block barrier per frame.
#ifndef ASM_GF_N
#define ASM_GF_N 16
#endif
constexpr int ASM_GF = ASM_GF_N;                      // frames per CTA of the gather variant (smaller tile: more CTAs per SM)
static_assert(ASM_GF >= 1 && ASM_GF <= 32 && 32 % ASM_GF == 0, "ASM_GF_N must divide 32 (write-out deals 32 / ASM_GF lines per warp)");
constexpr int ASM_GPAD = ASM_GF + 1;
constexpr int ASM_G_WARPS = 8, ASM_G_THREADS = 32 * ASM_G_WARPS;   // the gather variant keeps 8 warps: thread = equation
constexpr int ASM_KMAX = ASM_MAX_EQ / ASM_G_THREADS;

__global__ void __launch_bounds__(ASM_G_THREADS) k_assemble_gather(AsmParams P) {
    extern __shared__ __align__(16) float sh[];
    const int plane = (3 * P.max_eq + 3) & ~3;
    float *stage = sh;                                              // [2][3 planes][plane]: the frame's values, planar
    float *g_sh0 = sh + 6 * plane;                                  // [2][max_eq][9]: corner vectors, double buffered
    float *t_sh = g_sh0 + 2 * P.max_eq * 9;                         // [rows*3][33]
    int *src_sh = reinterpret_cast<int *>(t_sh + P.max_rows * 3 * ASM_GPAD);   // [max_eq]
    uint16_t *inc_sh = reinterpret_cast<uint16_t *>(src_sh + P.max_eq);        // [3 * max_eq]: the block's row incidences (CSR payload)
    const int4 blk = P.blocks[blockIdx.x];
    const int n_eq = blk.y - blk.x, n_rows = blk.w - blk.z;
    const int tile = blockIdx.y;
    const int frame0 = tile * ASM_GF;
    const int nvalid = max(0, min(ASM_GF, P.n_frames - frame0));    // 0: a tile past the batch only zero-fills its lanes
    for (int e = threadIdx.x; e < n_eq; e += ASM_G_THREADS) src_sh[e] = P.eq_src_local[blk.x + e];
    // the CSR of the block's rows does not depend on the frame either: payload to shared memory, the thread's row range
    // to registers
    const int inc0 = P.row_ptr[blk.z];
    for (int q = inc0 + threadIdx.x; q < P.row_ptr[blk.w]; q += ASM_G_THREADS) inc_sh[q - inc0] = P.inc[q];
    const int my_row = ASM_G_THREADS - 1 - threadIdx.x;       // rows are dealt from the top thread ids down (see below)
    const int my_q0 = my_row < n_rows ? P.row_ptr[blk.z + my_row] - inc0 : 0, my_q1 = my_row < n_rows ? P.row_ptr[blk.z + my_row + 1] - inc0 : 0;
    __syncthreads();
    int src_k[ASM_KMAX];
    float4 m0_k[ASM_KMAX], m1_k[ASM_KMAX];
#pragma unroll
    for (int k = 0; k < ASM_KMAX; ++k) {
        const int e = threadIdx.x + k * ASM_G_THREADS;
        src_k[k] = -1;
        m0_k[k] = m1_k[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n_eq) {
            src_k[k] = src_sh[e];
            m0_k[k] = __ldg(P.eq_meta + (size_t)(blk.x + e) * 2);
            m1_k[k] = __ldg(P.eq_meta + (size_t)(blk.x + e) * 2 + 1);
        }
    }
    // consecutive threads copy consecutive (equation, component) values, so that a warp's copy touches the few
    // cache lines of three or four triangles instead of 32; one commit group per frame and thread
    // The (source, destination) offsets of the thread's copies do not depend on the frame: computed once.
    constexpr int ASM_GCOPIES = (ASM_MAX_EQ * 9 + ASM_G_THREADS - 1) / ASM_G_THREADS;
    int g_src[ASM_GCOPIES], g_dst[ASM_GCOPIES];
#pragma unroll
    for (int i = 0; i < ASM_GCOPIES; ++i) {
        const int v = threadIdx.x + i * ASM_G_THREADS;
        g_src[i] = -1; g_dst[i] = 0;
        if (v < n_eq * 9) {
            const int e = v / 9, j = v - 9 * e, sr = src_sh[e];
            if (sr >= 0) { g_src[i] = sr * 9 + j; g_dst[i] = (j / 3) * plane + e * 3 + (j % 3); }
        }
    }
    auto gather = [&](int ft, int si) {
        const float *row = P.dgrad + (long long)(frame0 + ft) * P.frame_stride;
        const uint32_t dst = smem_u32(stage + si * 3 * plane);
#pragma unroll
        for (int i = 0; i < ASM_GCOPIES; ++i)
            if (g_src[i] >= 0)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * g_dst[i]), "l"(row + g_src[i]) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nvalid > 0) gather(0, 0);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    if (nvalid > 1) gather(1, 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");          // frame 0 has landed (this thread's part)
    __syncthreads();
    // iteration f computes the corner vectors of frame f into g_sh[f&1] and sums the rows of frame f-1
    for (int f = 0; f <= nvalid; ++f) {
        const float *st = stage + (f & 1) * 3 * plane;
        float *g_sh = g_sh0 + (f & 1) * P.max_eq * 9;
#pragma unroll
        for (int k = 0; k < ASM_KMAX; ++k) {
            const int e = threadIdx.x + k * ASM_G_THREADS;
            if (e >= n_eq || f >= nvalid) break;
            const int src = src_k[k];
            const float u0[3] = {m0_k[k].x, m0_k[k].y, m0_k[k].z}, u1[3] = {m0_k[k].w, m1_k[k].x, m1_k[k].y};
            float g2[3] = {0.f, 0.f, 0.f}, g3[3] = {0.f, 0.f, 0.f};    // src == -1: identity block, T - I = 0
            if (src >= 0) {
                float d[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) d[j] = st[(j / 3) * plane + e * 3 + (j % 3)];
                eq_vectors(P.mode, d, u0, u1, g2, g3);
            } else if (src == -2) {             // block left at zero by setZero (impl.hpp:224): T = 0, E = -I
#pragma unroll
                for (int c = 0; c < 3; ++c) { g2[c] = -u0[c]; g3[c] = -u1[c]; }
            }
            float *g = g_sh + e * 9;
#pragma unroll
            for (int c = 0; c < 3; ++c) { g[c] = -(g2[c] + g3[c]); g[3 + c] = g2[c]; g[6 + c] = g3[c]; }
        }
        if (f >= 1) {
            const float *gp = g_sh0 + ((f - 1) & 1) * P.max_eq * 9;
            // rows are dealt from the top thread ids down: the low threads carry the extra equations above
            for (int r = my_row; r < n_rows; r += ASM_G_THREADS) {
                const bool first = r == my_row;
                const int q0 = first ? my_q0 : P.row_ptr[blk.z + r] - inc0, q1 = first ? my_q1 : P.row_ptr[blk.z + r + 1] - inc0;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                for (int q = q0; q < q1; ++q) {
                    const float *g = gp + 3 * (int)inc_sh[q];
                    s0 += g[0]; s1 += g[1]; s2 += g[2];
                }
                float *t = t_sh + (3 * r) * ASM_GPAD + (f - 1);
                t[0] = s0; t[ASM_GPAD] = s1; t[2 * ASM_GPAD] = s2;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");       // frame f+1 has landed (this thread's copies; the barrier covers the rest)
        __syncthreads();
        if (f + 2 < nvalid) gather(f + 2, f & 1);                    // stage f&1 has been consumed
    }
    // write-out: a warp stores 32 / ASM_GF lines of ASM_GF frames per instruction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int LPW = 32 / ASM_GF;
    const int fl = lane % ASM_GF, sub = lane / ASM_GF;
    const int fr = frame0 + fl;
    float *dst_tile = P.rhs + (long long)(fr / P.L.FL) * P.L.tile_stride + fr % P.L.FL;
    for (int line = warp * LPW + sub; line < n_rows * 3; line += ASM_G_WARPS * LPW) {
        const int r = line / 3, c = line - 3 * r;
        dst_tile[(long long)P.row_perm[blk.z + r] * P.L.row_stride + c * P.L.c_stride] = fl < nvalid ? t_sh[line * ASM_GPAD + fl] : 0.f;
    }
}

// Gather variant, second generation: the same walk, colours and packed two-frames-per-lane transform as k_assemble, fed
// straight from a [frame][triangle][9] tensor.  An equation's input is one 36-byte record per frame; what is copied is the
// 16-byte aligned 48-byte span around it (rows are 16-byte aligned: frame_stride % 4 == 0, checked by the launcher).  A
// warp keeps the spans of its NEXT TWO equations in flight as 16-byte cp.async copies (six per lane and equation, no
// registers held; chunk q = lane + 32 k of the [frame][3 chunks] image, so consecutive lanes copy consecutive chunks of a
// span) into a private two-stage ring, and frees a stage as soon as its lane has pulled its two frames' spans into
// registers with six LDS.128 (conflict free: 3 l + c is distinct mod 8 over a quarter warp).  The lane's two frames are
// l and l + 32 -- position 2l / 2l + 1 of an accumulator line -- and the write-out undoes that pairing.
// One CTA per SM: accumulator 86 KB + rings 96 KB, i.e. 96 KB of loads in flight per SM (three stages when the row blocks are small enough).
constexpr int AG_REC = 12 * COMPACT_TILE;            // floats per image: [frame][12]
constexpr int AG_COPIES = AG_REC / 4 / 32;           // 16-byte copies per lane and equation

template <int SHIFT>
__device__ __forceinline__ void ag_select(const float4 (&a)[3], const float4 (&b)[3], float2 (&d)[9]) {
    const float wa[12] = {a[0].x, a[0].y, a[0].z, a[0].w, a[1].x, a[1].y, a[1].z, a[1].w, a[2].x, a[2].y, a[2].z, a[2].w};
    const float wb[12] = {b[0].x, b[0].y, b[0].z, b[0].w, b[1].x, b[1].y, b[1].z, b[1].w, b[2].x, b[2].y, b[2].z, b[2].w};
#pragma unroll
    for (int j = 0; j < 9; ++j) d[j] = make_float2(wa[SHIFT + j], wb[SHIFT + j]);
}

template <int AG_STAGES>
__global__ void __launch_bounds__(ASM_THREADS, 1) k_assemble_gather2(AsmParams P) {
    extern __shared__ __align__(16) float acc[];                     // [row][3][64], frame pairing (l, l + 32)
    constexpr int CT = COMPACT_TILE;
    const int4 blk = P.blocks[blockIdx.x];
    const int n_rows = blk.w - blk.z;
    const int frame0 = blockIdx.y * CT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int acc_floats = (P.max_rows * 3 * CT + 3) & ~3;
    float *ring = acc + acc_floats + warp * (AG_STAGES * AG_REC);
    int4 *walk_sh = reinterpret_cast<int4 *>(acc + acc_floats + ASM_WARPS * (AG_STAGES * AG_REC));
    for (int i = threadIdx.x; i < n_rows * 3 * CT; i += ASM_THREADS) acc[i] = 0.f;
    for (int i = lane; i < AG_STAGES * AG_REC; i += 32) ring[i] = 0.f;         // frames past the batch read zeros
    {
        const int w0 = P.warp_ptr[blockIdx.x * ASM_WARPS], w1 = P.warp_ptr[blockIdx.x * ASM_WARPS + ASM_WARPS];
        for (int i = threadIdx.x; i < w1 - w0; i += ASM_THREADS) walk_sh[i] = P.walk[w0 + i];
    }
    const int4 *walk = walk_sh + (P.warp_ptr[blockIdx.x * ASM_WARPS + warp] - P.warp_ptr[blockIdx.x * ASM_WARPS]);
    // chunk q = lane + 32 k of an image is chunk q % 3 of frame q / 3: its offset inside the tile's rows, once per lane
    int off[AG_COPIES];
    uint32_t valid = 0;
#pragma unroll
    for (int k = 0; k < AG_COPIES; ++k) {
        const int q = lane + 32 * k, f = q / 3;
        off[k] = (int)(f * P.frame_stride) + 4 * (q - 3 * f);
        if (frame0 + f < P.n_frames) valid |= 1u << k;
    }
    const float *tile_rows = P.dgrad + (long long)frame0 * P.frame_stride;
    const uint32_t ring_u32 = smem_u32(ring) + 16u * lane;
    __syncthreads();                                                 // accumulator and rings zeroed, walks in shared memory
    const int4 *wi = walk;                                           // issue pointer: AG_STAGES entries ahead of the transform
    auto issue = [&](int stage) {
        const int4 ent = *wi;
        if (ent.x != ASM_SCHED_END) ++wi;
        if (ent.x >= 0 && ent.y >= 0) {
            const float *src = tile_rows + ((long long)ent.y * 9 & ~3LL);
            const uint32_t dst = ring_u32 + 4u * AG_REC * (uint32_t)stage;
#pragma unroll
            for (int k = 0; k < AG_COPIES; ++k)
                if (valid >> k & 1u)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u * k), "l"(src + off[k]) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < AG_STAGES; ++i) issue(i);
    // the equation record (U0, U1, corner rows) of the NEXT entry is fetched one iteration ahead, under this entry's transform
    auto meta = [&](const int4 e, float4 &a, float4 &b) {
        if (e.x >= 0) { a = __ldg(P.eq_meta + (size_t)(blk.x + e.x) * 2); b = __ldg(P.eq_meta + (size_t)(blk.x + e.x) * 2 + 1); }
    };
    float4 m0n = make_float4(0.f, 0.f, 0.f, 0.f), m1n = m0n;
    meta(*walk, m0n, m1n);
    for (int n = 0, stage = 0;; ++n, stage = stage + 1 == AG_STAGES ? 0 : stage + 1) {
        const int4 ent = *walk++;
        const float4 m0 = m0n, m1 = m1n;
        if (ent.x != ASM_SCHED_END) meta(*walk, m0n, m1n);
        asm volatile("cp.async.wait_group %0;" ::"n"(AG_STAGES - 1) : "memory");   // this lane's copies of entry n have landed ...
        __syncwarp();                                                // ... and so have the other lanes'
        const bool rec = ent.x >= 0 && ent.y >= 0;
        float4 wa[3], wb[3];
        if (rec) {
            const float4 *st = reinterpret_cast<const float4 *>(ring + stage * AG_REC);
#pragma unroll
            for (int c = 0; c < 3; ++c) { wa[c] = st[3 * lane + c]; wb[c] = st[3 * (lane + 32) + c]; }
        }
        __syncwarp();                                                // the stage is free: refill it with entry n + AG_STAGES
        issue(stage);
        if (ent.x == ASM_SCHED_END) break;
        if (ent.x == ASM_SCHED_BARRIER) { __syncthreads(); continue; }
        float2 d[9];
        if (rec) {
            switch ((ent.y * 9) & 3) {
                case 0: ag_select<0>(wa, wb, d); break;
                case 1: ag_select<1>(wa, wb, d); break;
                case 2: ag_select<2>(wa, wb, d); break;
                default: ag_select<3>(wa, wb, d); break;
            }
        }
        eq_apply2(acc, lane, P.mode, ent.y, d, m0, m1);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // write-out: position 2l of a line is frame l, 2l + 1 is frame l + 32
    const int fa = frame0 + lane, fb = fa + 32;
    float *dst_a = P.rhs + (long long)(fa / P.L.FL) * P.L.tile_stride + fa % P.L.FL;
    float *dst_b = P.rhs + (long long)(fb / P.L.FL) * P.L.tile_stride + fb % P.L.FL;
    const float2 *acc2 = reinterpret_cast<const float2 *>(acc) + lane;
    for (int line = warp; line < n_rows * 3; line += ASM_WARPS) {
        const int r = line / 3, c = line - 3 * r;
        const float2 v = acc2[line * (CT / 2)];
        const long long o = (long long)P.row_perm[blk.z + r] * P.L.row_stride + c * P.L.c_stride;
        dst_a[o] = fa < P.n_frames ? v.x : 0.f;
        dst_b[o] = fb < P.n_frames ? v.y : 0.f;
    }
}

// Gather variant, third generation: k_assemble_gather2 with the copies handed to the TMA engine.  The dgrad tensor is a 2-D
// tensor map [frame][floats of a row]; an equation's 64 spans (one per frame of the tile, 48 bytes each, 359 KB apart) are
// ONE cp.async.bulk.tensor box {12 floats, 64 frames} issued by one lane into the warp's ring stage, completion on a
// per-stage mbarrier (frames past the batch are zero-filled by the engine).  The stage image, the conflict-free LDS.128
// pull and everything behind it are the second generation's.
__device__ __forceinline__ void ag_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

template <int AG_STAGES>
__global__ void __launch_bounds__(ASM_THREADS, 1) k_assemble_gather3(AsmParams P, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(16) float acc[];                     // [row][3][64], frame pairing (l, l + 32)
    constexpr int CT = COMPACT_TILE;
    const int4 blk = P.blocks[blockIdx.x];
    const int n_rows = blk.w - blk.z;
    const int frame0 = blockIdx.y * CT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int acc_floats = (P.max_rows * 3 * CT + 31) & ~31;
    // ring stages are TMA destinations: 128-byte aligned whatever the dynamic shared memory's own base is
    float *ring0 = acc + acc_floats + ((128u - (smem_u32(acc + acc_floats) & 127u)) & 127u) / 4u;
    float *ring = ring0 + warp * (AG_STAGES * AG_REC);
    int4 *walk_sh = reinterpret_cast<int4 *>(ring0 + ASM_WARPS * (AG_STAGES * AG_REC));
    uint64_t *bars = reinterpret_cast<uint64_t *>(walk_sh + P.max_walk * ASM_WARPS) + warp * AG_STAGES;
    for (int i = threadIdx.x; i < n_rows * 3 * CT; i += ASM_THREADS) acc[i] = 0.f;
    {
        const int w0 = P.warp_ptr[blockIdx.x * ASM_WARPS], w1 = P.warp_ptr[blockIdx.x * ASM_WARPS + ASM_WARPS];
        for (int i = threadIdx.x; i < w1 - w0; i += ASM_THREADS) walk_sh[i] = P.walk[w0 + i];
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < AG_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int4 *walk = walk_sh + (P.warp_ptr[blockIdx.x * ASM_WARPS + warp] - P.warp_ptr[blockIdx.x * ASM_WARPS]);
    const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
    __syncthreads();                                                 // accumulator zeroed, walks in shared memory, barriers live
    const int4 *wi = walk;                                           // issue pointer: AG_STAGES entries ahead of the transform
    auto issue = [&](int stage) {
        const int4 ent = *wi;
        if (ent.x != ASM_SCHED_END) ++wi;
        if (ent.x >= 0 && ent.y >= 0 && lane == 0) {
            const uint32_t bar = bars_u32 + 8u * (uint32_t)stage, dst = ring_u32 + 4u * AG_REC * (uint32_t)stage;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * AG_REC) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tmap), "r"((ent.y * 9) & ~3), "r"(frame0), "r"(bar) : "memory");
        }
    };
#pragma unroll
    for (int i = 0; i < AG_STAGES; ++i) issue(i);
    auto meta = [&](const int4 e, float4 &a, float4 &b) {
        if (e.x >= 0) { a = __ldg(P.eq_meta + (size_t)(blk.x + e.x) * 2); b = __ldg(P.eq_meta + (size_t)(blk.x + e.x) * 2 + 1); }
    };
    float4 m0n = make_float4(0.f, 0.f, 0.f, 0.f), m1n = m0n;
    meta(*walk, m0n, m1n);
    uint32_t phase = 0;                                              // bit s: parity of stage s's next completion
    for (int n = 0, stage = 0;; ++n, stage = stage + 1 == AG_STAGES ? 0 : stage + 1) {
        const int4 ent = *walk++;
        const float4 m0 = m0n, m1 = m1n;
        if (ent.x != ASM_SCHED_END) meta(*walk, m0n, m1n);
        const bool rec = ent.x >= 0 && ent.y >= 0;
        float4 wa[3], wb[3];
        if (rec) {
            ag_mbar_wait(bars_u32 + 8u * (uint32_t)stage, (phase >> stage) & 1u);   // the box of entry n has landed
            phase ^= 1u << stage;
            const float4 *st = reinterpret_cast<const float4 *>(ring + stage * AG_REC);
#pragma unroll
            for (int c = 0; c < 3; ++c) { wa[c] = st[3 * lane + c]; wb[c] = st[3 * (lane + 32) + c]; }
        }
        __syncwarp();                                                // the stage is free: refill it with entry n + AG_STAGES
        issue(stage);
        if (ent.x == ASM_SCHED_END) break;
        if (ent.x == ASM_SCHED_BARRIER) { __syncthreads(); continue; }
        float2 d[9];
        if (rec) {
            switch ((ent.y * 9) & 3) {
                case 0: ag_select<0>(wa, wb, d); break;
                case 1: ag_select<1>(wa, wb, d); break;
                case 2: ag_select<2>(wa, wb, d); break;
                default: ag_select<3>(wa, wb, d); break;
            }
        }
        eq_apply2(acc, lane, P.mode, ent.y, d, m0, m1);
    }
    __syncthreads();
    // write-out: position 2l of a line is frame l, 2l + 1 is frame l + 32
    const int fa = frame0 + lane, fb = fa + 32;
    float *dst_a = P.rhs + (long long)(fa / P.L.FL) * P.L.tile_stride + fa % P.L.FL;
    float *dst_b = P.rhs + (long long)(fb / P.L.FL) * P.L.tile_stride + fb % P.L.FL;
    const float2 *acc2 = reinterpret_cast<const float2 *>(acc) + lane;
    for (int line = warp; line < n_rows * 3; line += ASM_WARPS) {
        const int r = line / 3, c = line - 3 * r;
        const float2 v = acc2[line * (CT / 2)];
        const long long o = (long long)P.row_perm[blk.z + r] * P.L.row_stride + c * P.L.c_stride;
        dst_a[o] = fa < P.n_frames ? v.x : 0.f;
        dst_b[o] = fb < P.n_frames ? v.y : 0.f;
    }
}

size_t assemble_gather3_smem(const DevicePlan &d, int stages) {
    return (size_t)((d.asm_max_rows * 3 * COMPACT_TILE + 31) & ~31) * sizeof(float) + (size_t)ASM_WARPS * stages * AG_REC * sizeof(float) +
           (size_t)d.asm_max_walk * ASM_WARPS * sizeof(int4) + (size_t)ASM_WARPS * stages * 8 + 128;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

size_t assemble_gather2_smem(const DevicePlan &d, int stages) {
    return (size_t)((d.asm_max_rows * 3 * COMPACT_TILE + 3) & ~3) * sizeof(float) + (size_t)ASM_WARPS * stages * AG_REC * sizeof(float) +
           (size_t)d.asm_max_walk * ASM_WARPS * sizeof(int4);
}

cudaError_t launch_assembly(const DevicePlan &d, const float *dgrad, long long frame_stride, bool staged,
                            int n_frames, int mode, float *rhs, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    AsmParams P{d.asm_blocks, d.asm_walk, d.asm_warp_ptr, d.asm_eq_meta, d.asm_row_perm, d.asm_eq_src_local, d.asm_row_ptr, d.asm_inc,
                d.asm_max_eq, dgrad, frame_stride, d.compact_s_rows, rhs, n_frames, mode, d.asm_max_rows, d.asm_max_walk, d.layout};
    dim3 grid((unsigned)d.n_asm_blocks, (unsigned)((n_frames + COMPACT_TILE - 1) / COMPACT_TILE));
    if (staged) {
        const size_t smem = (size_t)d.asm_max_rows * 3 * COMPACT_TILE * sizeof(float) + (size_t)d.asm_max_walk * ASM_WARPS * sizeof(int4);
        cudaError_t e = cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_assemble<<<grid, ASM_THREADS, smem, stream>>>(P);
    } else if ((d.asm_gather_gen >= 3 || (d.asm_gather_gen == 0 && frame_stride <= 262144)) && frame_stride % 4 == 0 &&
               reinterpret_cast<uintptr_t>(dgrad) % 16 == 0 && encode_tiled_fn() &&
               assemble_gather3_smem(d, 2) <= (size_t)227 * 1024) {
        // the dgrad tensor as a 2-D tensor map [frame][row floats], box = {12 floats, 64 frames}
        CUtensorMap tmap;
        const cuuint64_t dims[2] = {(cuuint64_t)frame_stride, (cuuint64_t)n_frames};
        const cuuint64_t strides[1] = {(cuuint64_t)frame_stride * 4};
        const cuuint32_t box[2] = {12, (cuuint32_t)COMPACT_TILE}, estr[2] = {1, 1};
        if (encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(dgrad), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
        const bool three = assemble_gather3_smem(d, 3) <= (size_t)227 * 1024;
        const size_t smem = assemble_gather3_smem(d, three ? 3 : 2);
        auto kern = three ? k_assemble_gather3<3> : k_assemble_gather3<2>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, ASM_THREADS, smem, stream>>>(P, tmap);
    } else if ((d.asm_gather_gen >= 2 || d.asm_gather_gen == 0) && (long long)(COMPACT_TILE - 1) * frame_stride + 12 < 0x7fffffffLL && frame_stride % 4 == 0 &&
               reinterpret_cast<uintptr_t>(dgrad) % 16 == 0) {   // 16-byte copies of aligned spans
        // as many stages (equations in flight per warp) as fit beside the accumulator
        const bool three = assemble_gather2_smem(d, 3) <= (size_t)227 * 1024;
        const size_t smem = assemble_gather2_smem(d, three ? 3 : 2);
        auto kern = three ? k_assemble_gather2<3> : k_assemble_gather2<2>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, ASM_THREADS, smem, stream>>>(P);
    } else {
        const size_t plane = (size_t)((3 * d.asm_max_eq + 3) & ~3);
        const size_t smem = (6 * plane + 2 * (size_t)d.asm_max_eq * 9 + (size_t)d.asm_max_rows * 3 * ASM_GPAD + d.asm_max_eq) * sizeof(float) +
                            (((size_t)d.asm_max_eq * 3 * sizeof(uint16_t) + 15) & ~(size_t)15);
        cudaError_t e = cudaFuncSetAttribute(k_assemble_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // whole 32-frame groups are covered so that the scratch's idle lanes of a partial group hold zeros
        grid.y = (unsigned)((n_frames + 31) / 32 * (32 / ASM_GF));
        k_assemble_gather<<<grid, ASM_G_THREADS, smem, stream>>>(P);
    }
    g_launches++;
    return cudaGetLastError();
}

// =============================================================================================
// K3: batched multi-RHS sparse triangular solve (forward + backward): level-scheduled supernodal sweeps
// over the program built by schedule.cpp.  One CTA = one tile of 32 frames (lane = frame, 3 coordinates
// per lane); persistent CTAs loop over tiles.  Warp roles:
//   warp 0   streamer : copies the task lists' stages global -> shared ring (cp.async.bulk + mbarrier)
//   warp 1   IO       : per phase, TMA bulk-loads the piece's rows scratch -> state slots (prefetched one
//                       phase ahead) and bulk-stores finished rows state -> scratch
//   warps 2+ consumers: run the row tasks of a level (dealt round-robin), one named barrier per level
// The factor is never re-read from HBM per frame: every task entry fetched from the ring is applied to
// 96 right-hand sides (32 frames x 3 coordinates) of the tile.
constexpr int RING = 3;
#ifndef SOLVE_NCW
#define SOLVE_NCW 12
#endif
constexpr int NCW = SOLVE_NCW;                  // consumer warps
constexpr int SOLVE_THREADS = 32 * (NCW + 2);

struct SolveParams {
    const uint8_t *prog;
    const uint32_t *stage_off;
    const IoDesc *io_desc;
    const IoPhase *io_phase;
    int n_stages, n_slots, n_phases_fwd, n_phases_bwd;
    float *scratch;                             // [n_tiles][n_free][3][32]: rhs in, x out (in place)
    int n_free, n_tiles;
    long long *prof;                            // optional [grid][8] cycle counters of consumer warp 0 (NULL = off)
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }

// One task on one warp (lane = frame).  Entry lists are padded by the host to whole batches (4 sources of
// kind B, 4 packed pairs of kind A), and the next batch's entries are fetched while the current batch's
// state values are in flight, so a batch costs one shared-memory round trip instead of two.
// With fewer than 32 frames per tile (large factors) the warp's 32 / F lane groups split a task's entry batches among
// them and add their partial sums with shuffles, instead of mirroring each other.
template <int COORD_STRIDE>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
    for (int off = COORD_STRIDE; off < 32; off <<= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

template <int COORD_STRIDE>
__device__ __forceinline__ void run_row_task(const uint8_t *task, uint8_t *state_lane) {
    constexpr int G = 32 / COORD_STRIDE;                            // lane groups
    const int grp = (threadIdx.x & 31) / COORD_STRIDE;
    const uint4 th = *reinterpret_cast<const uint4 *>(task);       // TaskHeader
    const int n = (int)(th.w & 0xFFFFFFu);
    const uint4 *e = reinterpret_cast<const uint4 *>(task + 16);
    if (th.w & TASK_GROUP) {
        // kind B: up to three target rows share the sources -> 9 FMAs per 3 shared-memory value loads
        float a[3][3] = {};
        constexpr int BB = TASK_BATCH_B;
        uint4 p[BB];
        const int k0 = grp * BB;
#pragma unroll
        for (int j = 0; j < BB; ++j) p[j] = e[min(k0, n - BB) + j];
        for (int k = k0; k < n; k += G * BB) {
            float v[BB][3];
#pragma unroll
            for (int j = 0; j < BB; ++j) {
                const float *s = reinterpret_cast<const float *>(state_lane + p[j].x);
                v[j][0] = s[0]; v[j][1] = s[COORD_STRIDE]; v[j][2] = s[2 * COORD_STRIDE];
            }
            uint4 q[BB];
            const int kn = (k + G * BB < n) ? k + G * BB : k;       // last batch re-reads itself (harmless)
#pragma unroll
            for (int j = 0; j < BB; ++j) q[j] = e[kn + j];
#pragma unroll
            for (int j = 0; j < BB; ++j) {
                const float c0 = __uint_as_float(p[j].y), c1 = __uint_as_float(p[j].z), c2 = __uint_as_float(p[j].w);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    a[0][d] = fmaf(c0, v[j][d], a[0][d]);
                    a[1][d] = fmaf(c1, v[j][d], a[1][d]);
                    a[2][d] = fmaf(c2, v[j][d], a[2][d]);
                }
            }
#pragma unroll
            for (int j = 0; j < BB; ++j) p[j] = q[j];
        }
        if (G > 1) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int d = 0; d < 3; ++d) a[r][d] = group_sum<COORD_STRIDE>(a[r][d]);
        }
        const uint32_t tg[3] = {th.x, th.y, th.z};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (tg[r] == 0xFFFFFFFFu) continue;
            float *t = reinterpret_cast<float *>(state_lane + tg[r]);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (!(th.w & (TASK_OVERWRITE << r))) { v0 = t[0]; v1 = t[COORD_STRIDE]; v2 = t[2 * COORD_STRIDE]; }
            t[0] = v0 - a[r][0];
            t[COORD_STRIDE] = v1 - a[r][1];
            t[2 * COORD_STRIDE] = v2 - a[r][2];
        }
        return;
    }
    // kind A: one target row, two entries {coeff, src} per uint4, 4 uint4 (8 entries) per batch
    float a[3] = {}, b[3] = {};
    const int nq = n >> 1;                                          // number of uint4 pairs, multiple of 4
    const int q0 = grp * 4, qs = min(q0, nq - 4);
    uint4 p[4] = {e[qs], e[qs + 1], e[qs + 2], e[qs + 3]};
    for (int k = q0; k < nq; k += G * 4) {
        float v[4][2][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *s0 = reinterpret_cast<const float *>(state_lane + p[j].y);
            const float *s1 = reinterpret_cast<const float *>(state_lane + p[j].w);
            v[j][0][0] = s0[0]; v[j][0][1] = s0[COORD_STRIDE]; v[j][0][2] = s0[2 * COORD_STRIDE];
            v[j][1][0] = s1[0]; v[j][1][1] = s1[COORD_STRIDE]; v[j][1][2] = s1[2 * COORD_STRIDE];
        }
        uint4 q[4];
        const int kn = (k + G * 4 < nq) ? k + G * 4 : k;
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = e[kn + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float c0 = __uint_as_float(p[j].x), c1 = __uint_as_float(p[j].z);
#pragma unroll
            for (int d = 0; d < 3; ++d) { a[d] = fmaf(c0, v[j][0][d], a[d]); b[d] = fmaf(c1, v[j][1][d], b[d]); }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = q[j];
    }
    if (G > 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { a[d] = group_sum<COORD_STRIDE>(a[d] + b[d]); b[d] = 0.f; }
    }
    float *t = reinterpret_cast<float *>(state_lane + th.x);
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (!(th.w & TASK_OVERWRITE)) { v0 = t[0]; v1 = t[COORD_STRIDE]; v2 = t[2 * COORD_STRIDE]; }
    t[0] = v0 - (a[0] + b[0]);
    t[COORD_STRIDE] = v1 - (a[1] + b[1]);
    t[2 * COORD_STRIDE] = v2 - (a[2] + b[2]);
}

// shared memory map: [ring RING*STAGE_BYTES][barriers 8*(2*RING+4)][pad to 128][state n_slots*384]
constexpr int SOLVE_BAR_BYTES = 128;
constexpr int BAR_FULL = 0, BAR_EMPTY = RING, BAR_LD = 2 * RING, BAR_DONE = 2 * RING + 2;

template <int F>
__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(SolveParams P) {
    constexpr int SLOT_WORDS = 3 * F, SLOT_BYTES = 12 * F;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RING * STAGE_BYTES);
    uint8_t *state = smem + RING * STAGE_BYTES + SOLVE_BAR_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) {
            mbar_init(smem_u32(&bars[BAR_FULL + s]), 1);        // streamer's arrive.expect_tx
            mbar_init(smem_u32(&bars[BAR_EMPTY + s]), NCW);     // one arrive per consumer warp
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars[BAR_LD + s]), 1);          // IO warp's arrive.expect_tx
            mbar_init(smem_u32(&bars[BAR_DONE + s]), NCW);      // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    const int n_phases = P.n_phases_fwd + P.n_phases_bwd;

    if (warp == 0) {
        // ------------------------------------------------------------------ streamer
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int s = 0; s < P.n_stages; ++s, ++it) {
                    const uint32_t slot = it % RING, phase = (it / RING) & 1u;
                    mbar_wait(smem_u32(&bars[BAR_EMPTY + slot]), phase ^ 1u);
                    const uint32_t off = P.stage_off[s], bytes = P.stage_off[s + 1] - off;
                    const uint32_t full = smem_u32(&bars[BAR_FULL + slot]);
                    mbar_arrive_expect_tx(full, bytes);
                    tma_bulk_g2s(smem_u32(ring + slot * STAGE_BYTES), P.prog + off, bytes, full);
                }
            }
        }
        return;
    }
    if (warp == 1) {
        // ------------------------------------------------------------------ IO
        if (lane == 0) {
            uint32_t g = 0;                                     // running phase counter (consumers count alike)
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                uint8_t *tile_base = reinterpret_cast<uint8_t *>(P.scratch + (long long)tile * P.n_free * SLOT_WORDS);
                auto issue_loads = [&](int q, uint32_t gq) {
                    const IoPhase ph = P.io_phase[q];
                    uint32_t bytes = 0;
                    for (uint32_t i = ph.load_begin; i < ph.load_end; ++i) bytes += P.io_desc[i].n_rows * SLOT_BYTES;
                    const uint32_t bar = smem_u32(&bars[BAR_LD + (gq & 1u)]);
                    mbar_arrive_expect_tx(bar, bytes);
                    for (uint32_t i = ph.load_begin; i < ph.load_end; ++i) {
                        const IoDesc d = P.io_desc[i];
                        tma_bulk_g2s(smem_u32(state + (size_t)d.slot * SLOT_BYTES), tile_base + (size_t)d.row * SLOT_BYTES,
                                     d.n_rows * SLOT_BYTES, bar);
                    }
                };
                int q0 = 0;
                for (int sweep = 0; sweep < 2; ++sweep) {
                    const int nq = sweep == 0 ? P.n_phases_fwd : P.n_phases_bwd;
                    issue_loads(q0, g);
                    if (nq > 1) issue_loads(q0 + 1, g + 1);
                    for (int q = 0; q < nq; ++q, ++g) {
                        mbar_wait(smem_u32(&bars[BAR_DONE + (g & 1u)]), (g >> 1) & 1u);   // consumers finished phase q
                        const IoPhase ph = P.io_phase[q0 + q];
                        for (uint32_t i = ph.store_begin; i < ph.store_end; ++i) {
                            const IoDesc d = P.io_desc[i];
                            tma_bulk_s2g(tile_base + (size_t)d.row * SLOT_BYTES, smem_u32(state + (size_t)d.slot * SLOT_BYTES),
                                         d.n_rows * SLOT_BYTES);
                        }
                        tma_commit();
                        if (q + 2 < nq) {
                            tma_wait_read0();                   // the stored slots may be reused by these loads
                            issue_loads(q0 + q + 2, g + 2);
                        }
                    }
                    tma_wait_all0();                            // the next sweep / tile re-reads what was stored
                    q0 += nq;
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    const int cw = warp - 2;
    uint8_t *state_lane = state + (lane % F) * 4;     // F < 32: the upper lanes mirror the lower ones
    uint32_t it = 0, g = 0;
    (void)n_phases;
    const bool prof = P.prof != nullptr && cw == 0;
    long long c_stage = 0, c_ld = 0, c_task = 0, c_bar = 0, c_t0 = prof ? clock64() : 0, c_a = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (int s = 0; s < P.n_stages; ++s, ++it) {
            const uint32_t slot = it % RING, phase = (it / RING) & 1u;
            if (prof) c_a = clock64();
            mbar_wait(smem_u32(&bars[BAR_FULL + slot]), phase);
            if (prof) c_stage += clock64() - c_a;
            const uint8_t *stage = ring + slot * STAGE_BYTES;
            const int n_ops = (int)reinterpret_cast<const uint32_t *>(stage)[0];
            uint32_t at = 16;
            for (int o = 0; o < n_ops; ++o) {
                const uint4 hw = *reinterpret_cast<const uint4 *>(stage + at);   // OpHeader
                const uint32_t type = hw.x & 0xFFFFu, flags = hw.x >> 16;
                if (type == OP_ROWS) {
                    const uint32_t *table = reinterpret_cast<const uint32_t *>(stage + hw.z);
                    if (prof) c_a = clock64();
                    for (uint32_t t = cw; t < hw.y; t += NCW) run_row_task<F>(stage + table[t], state_lane);
                    if (prof) { long long c_b = clock64(); c_task += c_b - c_a; c_a = c_b; }
                    at = hw.w;
                    if (flags & OPF_SYNC_AFTER) consumer_bar();
                    if (prof) c_bar += clock64() - c_a;
                } else if (type == OP_PHASE_BEGIN) {
                    if (prof) c_a = clock64();
                    mbar_wait(smem_u32(&bars[BAR_LD + (g & 1u)]), (g >> 1) & 1u);
                    if (prof) c_ld += clock64() - c_a;
                    at += 16;
                } else {   // OP_PHASE_END: the level barrier before it has ordered all consumer writes
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_DONE + (g & 1u)]));
                    ++g;
                    at += 16;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_EMPTY + slot]));
        }
    }
    if (prof && lane == 0) {
        long long *o = P.prof + (long long)blockIdx.x * 8;
        o[0] = clock64() - c_t0; o[1] = c_stage; o[2] = c_ld; o[3] = c_task; o[4] = c_bar;
    }
}

size_t scratch_floats(const DevicePlan &d, int n_frames) {
    // K2 writes whole tiles of COMPACT_TILE frames
    const size_t nt = ((size_t)n_frames + COMPACT_TILE - 1) / COMPACT_TILE * COMPACT_TILE;
    return (nt + d.layout.FL - 1) / d.layout.FL * (size_t)d.layout.tile_stride;
}

size_t solve_smem_bytes(int n_slots, int frames_per_tile) {
    return (size_t)RING * STAGE_BYTES + SOLVE_BAR_BYTES + (size_t)n_slots * slot_bytes(frames_per_tile);
}

template <int F>
static cudaError_t configure_solve_f(DevicePlan &d) {
    const size_t smem = solve_smem_bytes(d.n_slots, F);
    cudaError_t e = cudaFuncSetAttribute(k_solve<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_solve<F>, SOLVE_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    d.solve_ctas_per_sm = n;
    return cudaSuccess;
}

cudaError_t configure_solve(DevicePlan &d) {
    switch (d.frames_per_tile) {
        case 32: return configure_solve_f<32>(d);
        case 16: return configure_solve_f<16>(d);
        case 8: return configure_solve_f<8>(d);
        default: return cudaErrorInvalidValue;
    }
}

template <int F>
static cudaError_t launch_solve_f(const DevicePlan &d, float *scratch, int n_frames, cudaStream_t stream) {
    const size_t smem = solve_smem_bytes(d.n_slots, F);
    // the attribute is per function and device, not per handle: another handle may have set a smaller size since
    cudaError_t e = cudaFuncSetAttribute(k_solve<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int n_tiles = (n_frames + F - 1) / F;
    SolveParams P{d.prog, d.stage_off, d.io_desc, d.io_phase, d.n_stages, d.n_slots, d.n_phases_fwd, d.n_phases_bwd,
                  scratch, d.n_free, n_tiles, d.solve_prof};
    int grid = d.sm_count * d.solve_ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    k_solve<F><<<grid, SOLVE_THREADS, smem, stream>>>(P);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_solve(const DevicePlan &d, float *scratch, int n_frames, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    switch (d.frames_per_tile) {
        case 32: return launch_solve_f<32>(d, scratch, n_frames, stream);
        case 16: return launch_solve_f<16>(d, scratch, n_frames, stream);
        case 8: return launch_solve_f<8>(d, scratch, n_frames, stream);
        default: return cudaErrorInvalidValue;
    }
}

// =============================================================================================
// K5: output.  Transposes the solved displacement rows (one line of FR consecutive frames per free vertex
// coordinate in the solve scratch) into the reference layout out[frame][vertex][3] (pybind.cpp:108), adds the
// fp64-computed base solution (kept as a hi/lo float pair) and copies the constrained vertices through
// (impl.hpp:295-308) -- every output byte is written exactly once, coalesced.
// A CTA owns 64 vertices x FR frames: phase 1 stages the chunk's free lines (+ base) in shared memory, one
// coalesced line per warp instruction; phase 2 has one thread per output element of the chunk walk the frames,
// its line index / constant in registers, so a frame's 192 floats leave as six full-warp stores.  The per-chunk
// tables (which element is which line, constants, base) are built on the host (api.cpp: upload_base).
constexpr int OUT_VC = 64;                       // vertices per CTA
constexpr int OUT_THREADS = OUT_VC * 3;
struct OutParams {
    const float *scratch;
    const int16_t *line_of;                      // [chunks][192] free line of the element inside its chunk, -1 = constrained
    const float *cval;                           // [chunks][192] constrained value of the element
    const int32_t *line_ptr;                     // [chunks + 1] first line of the chunk
    const int32_t *line_off;                     // [lines] scratch offset of the line inside a tile
    const float *line_hi, *line_lo;              // [lines] base solution
    float *out;
    int n_frames, n_verts, FR, FC;               // FR = frames per scratch line, FC <= FR = frames per CTA
    long long tile_stride;
};

constexpr int OUT_L2_AHEAD = 16;

__global__ void __launch_bounds__(OUT_THREADS) k_output(OutParams P) {
    extern __shared__ float t_sh[];              // [line][FC + 1]
    const int chunk = blockIdx.x, grp = blockIdx.y;
    const int frame0 = grp * P.FC;
    const int nvalid = min(P.FC, P.n_frames - frame0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l0 = P.line_ptr[chunk], nl = P.line_ptr[chunk + 1] - l0;
    const int pitch = P.FC + 1;
    const float *src_tile = P.scratch + (long long)(frame0 / P.FR) * P.tile_stride + frame0 % P.FR;
    {   // pull the lines of the CTA OUT_L2_AHEAD frame groups further on (same chunk: about one wave of CTAs later) into
        // L2, one 128-byte line per thread: its staging loop then waits for L2 instead of DRAM (1.50 -> 1.44 ms; 6 ... 32
        // groups ahead measure the same)
        const int f2 = frame0 + OUT_L2_AHEAD * P.FC;
        if (f2 < P.n_frames) {
            const float *t2 = P.scratch + (long long)(f2 / P.FR) * P.tile_stride + f2 % P.FR;
            for (int i = threadIdx.x; i < nl; i += OUT_THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(t2 + P.line_off[l0 + i]));
        }
    }
    if (lane < P.FC) {
#pragma unroll 4
        for (int i = warp; i < nl; i += OUT_THREADS / 32)
            t_sh[i * pitch + lane] = P.line_hi[l0 + i] + (P.line_lo[l0 + i] + __ldcs(src_tile + P.line_off[l0 + i] + lane));
    }
    __syncthreads();
    const int e = threadIdx.x;
    const long long row = (long long)P.n_verts * 3;
    if ((long long)chunk * OUT_THREADS + e >= row) return;
    const int line = P.line_of[chunk * OUT_THREADS + e];
    const float c = P.cval[chunk * OUT_THREADS + e];
    float *dst = P.out + (long long)frame0 * row + (long long)chunk * OUT_THREADS + e;
    if (line < 0) {
#pragma unroll 4
        for (int f = 0; f < nvalid; ++f) __stcs(dst + f * row, c);
    } else {
        const float *src = t_sh + line * pitch;
#pragma unroll 4
        for (int f = 0; f < nvalid; ++f) __stcs(dst + f * row, src[f]);
    }
}

// Second generation of the output kernel: no shared-memory staging and no block barrier.  A thread still owns one output
// element of the chunk for FC consecutive frames; if the element is a free coordinate its FC solved values are ONE
// contiguous run of the scratch line, so the thread pulls them with FC / 4 independent 16-byte loads issued up front (all
// of a CTA's loads are in flight at once; the sectors a warp's first load touches are completed by its next ones out of L1),
// adds the base, and every frame's store stays one full-warp instruction (constrained lanes select their constant).
// Measured equal to the staged kernel (1.41 against 1.43 ms per 75 600 FLAME frames): the kernel is bound by the store
// pattern itself -- tools/micro/outbw.cu writes the same 192-float x 32-frame pieces with no loads at all in 0.93 ms
// (4.9 TB/s), while whole frames written front to back reach 6 - 7 TB/s.  A frame-major variant (a CTA owns four whole
// frames, stages their free coordinates with one 16-byte load per scratch line) was tried and dropped: 2.3 - 3.0 ms, because
// a CTA then uses 16 bytes of every 512-byte scratch line it touches.
template <int FC>
__global__ void __launch_bounds__(OUT_THREADS) k_output2(OutParams P) {
    const int chunk = blockIdx.x, frame0 = blockIdx.y * FC;
    const int nvalid = min(FC, P.n_frames - frame0);
    const int e = threadIdx.x;
    const long long row = (long long)P.n_verts * 3;
    if ((long long)chunk * OUT_THREADS + e >= row) return;
    const int line = P.line_of[chunk * OUT_THREADS + e];
    float c = P.cval[chunk * OUT_THREADS + e];
    float4 v[FC / 4];
    float lo = 0.f;
    if (line >= 0) {
        const int l = P.line_ptr[chunk] + line;
        const float4 *src = reinterpret_cast<const float4 *>(P.scratch + (long long)(frame0 / P.FR) * P.tile_stride + frame0 % P.FR +
                                                             P.line_off[l]);
#pragma unroll
        for (int i = 0; i < FC / 4; ++i) v[i] = __ldcs(src + i);
        c = P.line_hi[l];
        lo = P.line_lo[l];
    } else {
#pragma unroll
        for (int i = 0; i < FC / 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float *dst = P.out + (long long)frame0 * row + (long long)chunk * OUT_THREADS + e;
    // c + (lo + x): for a constrained element lo = x = 0 and c is its constant, so the arithmetic is the same instruction
    // stream for every lane and the value of a free element is hi + (lo + x) as before
#pragma unroll
    for (int i = 0; i < FC / 4; ++i) {
        const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = 4 * i + j;
            if (f < nvalid) __stcs(dst + (long long)f * row, line >= 0 ? c + (lo + x[j]) : c);
        }
    }
}

cudaError_t launch_output(const DevicePlan &d, const DevicePlan::OutTables &t, const float *scratch, int n_frames, float *out,
                          cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int FR = d.layout.FL, FC = std::min(FR, d.out_gen >= 2 ? d.out_fc : 32);
    OutParams P{scratch, t.line_of, t.cval, t.line_ptr, t.line_off, t.line_hi, t.line_lo,
                out, n_frames, t.n_rows, FR, FC, d.layout.tile_stride};
    dim3 grid((unsigned)((t.n_rows + OUT_VC - 1) / OUT_VC), (unsigned)((n_frames + FC - 1) / FC));
    if (d.out_gen >= 2 && FR % FC == 0 && FC % 4 == 0) {
        switch (FC) {
            case 32: k_output2<32><<<grid, OUT_THREADS, 0, stream>>>(P); break;
            case 16: k_output2<16><<<grid, OUT_THREADS, 0, stream>>>(P); break;
            default: k_output2<8><<<grid, OUT_THREADS, 0, stream>>>(P); break;
        }
        g_launches++;
        return cudaGetLastError();
    }
    const size_t smem = (size_t)std::max(t.max_lines, 1) * (FC + 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_output, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_output<<<grid, OUT_THREADS, smem, stream>>>(P);
    g_launches++;
    return cudaGetLastError();
}

// Free rows [n_frames][n_free*3] -> all vertices [n_frames][n_verts*3]: the receiving side of a gather that moved only
// the free rows (the constrained rows are per-handle constants, impl.hpp:302-308).  thread = output element, EXP_FRAMES
// frames per CTA; a frame's 15 KB of free rows stay in L1 / L2 while its 60 KB leave as full-warp stores.
constexpr int EXP_THREADS = 256, EXP_FRAMES = 16;
__global__ void __launch_bounds__(EXP_THREADS) k_expand(const float *__restrict__ free_rows, const int32_t *__restrict__ src,
                                                        const float *__restrict__ cval, float *__restrict__ out, int n_frames,
                                                        int row_out, int row_in) {
    const int e = blockIdx.x * EXP_THREADS + threadIdx.x;
    if (e >= row_out) return;
    const int sidx = src[e];
    const float c = cval[e];
    const int f0 = blockIdx.y * EXP_FRAMES, f1 = min(n_frames, f0 + EXP_FRAMES);
    if (sidx < 0) {
        for (int f = f0; f < f1; ++f) __stcs(out + (long long)f * row_out + e, c);
    } else {
#pragma unroll 4
        for (int f = f0; f < f1; ++f) __stcs(out + (long long)f * row_out + e, __ldg(free_rows + (long long)f * row_in + sidx));
    }
}

cudaError_t launch_expand(const DevicePlan &d, const float *free_rows, int n_frames, float *out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int row_out = d.n_verts * 3, row_in = d.n_free * 3;
    dim3 grid((unsigned)((row_out + EXP_THREADS - 1) / EXP_THREADS), (unsigned)((n_frames + EXP_FRAMES - 1) / EXP_FRAMES));
    k_expand<<<grid, EXP_THREADS, 0, stream>>>(free_rows, d.exp_src, d.exp_cval, out, n_frames, row_out, row_in);
    g_launches++;
    return cudaGetLastError();
}

// =============================================================================================
// Full-layout decode (exact fp32 FMA on CUDA cores; not on the hot path -- the path's K1 is decode_tc.cu): dgrad[f][tri][0..5] = coeff_s[f] . Ws[tri*6+s] + ms,
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

cudaError_t launch_decode_full(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                               float *dgrad_out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int ntri = d.n_pca_tris;                 // rows of the basis sdfa_set_pca uploaded (not the template's)
    const long long stride = (long long)ntri * 9;
    dim3 gs((unsigned)((ntri * 6 + DEC_TJ - 1) / DEC_TJ), (unsigned)((n_frames + DEC_TF - 1) / DEC_TF));
    k_decode<<<gs, 256, 0, stream>>>(coeff_scale, d.k_scale, d.wfull_scale, d.mfull_scale, ntri * 6, 6, 0, n_frames, dgrad_out, stride);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 gr((unsigned)((ntri * 3 + DEC_TJ - 1) / DEC_TJ), (unsigned)((n_frames + DEC_TF - 1) / DEC_TF));
    k_decode<<<gr, 256, 0, stream>>>(coeff_rotat, d.k_rotat, d.wfull_rotat, d.mfull_rotat, ntri * 3, 3, 6, n_frames, dgrad_out, stride);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace sdfa
