// decode_tc.cu -- K1: PCA coefficients -> dgrad of the needed source triangles, on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged in shared memory by TMA
// bulk copies).  Replaces PcaInversion.forward x2 + the scale/rotation interleave of data_to_anime_feat
// (reference speech_anime/modules/output_module.py:115-116, speech_anime/model/model.py:246-257), which
// is a cuBLAS SGEMM + cat in the reference.
//
// GEMM shape (per basis: "scale" K=85 with 6 outputs per triangle, "rotation" K=180 with 3):
//      D[n, m] = sum_k X[n, k] * W[m, k]       n = frame, m = basis row (output value = "slot")
// i.e. the frames are the MMA M dimension (TMEM lanes) and the basis rows the N dimension (TMEM columns): an
// epilogue warp owns 32 consecutive frames (lane = frame), holds one output value per register, and every
// store instruction writes one full 128-byte line of the frame-tiled compact dgrad [tile of 64 frames][slot][64]
// (its half of the slot's 64 frames) -- the layout the assembly kernel (lane = two frames) reads back with coalesced
// 256-byte lines.  Two CTAs form a pair (cta_group::2): ONE instruction stream, issued by the pair's leader, drives the
// tensor cores of both SMs on a 256-frame x 256-row tile, each CTA staging its own 128 frames and half of the basis rows.
//
// fp32 accuracy from TF32 tensor cores: both operands are split x = hi + lo with hi = x truncated to
// TF32 (the 19 bits the tensor core reads) and lo = x - hi (exact), and every K-step issues three MMAs
//      W_hi X_hi + W_hi X_lo + W_lo X_hi                  ("3xTF32", error ~2^-21 per product)
// into the same fp32 accumulator.  The basis is split and pre-tiled once on the host (sdfa_set_pca); the
// coefficients are split per call by k_split_coeffs.  Tiles are stored in global memory as exact images
// of the 128-byte-swizzled K-major shared-memory layout the UMMA descriptors expect, so one
// cp.async.bulk per operand half and K-block fills a ring slot (no tensor maps needed).
#include "device_plan.hpp"

#include <cstring>
#include <vector>

namespace sdfa {


namespace {

constexpr int TC_BM = 256;          // basis rows per tile (UMMA N): 256 halves the frames-operand traffic per MMA
constexpr int TC_BN = 128;          // frames per tile (UMMA M = TMEM lanes)
constexpr int TC_BK = 32;           // floats per K-block = one 128-byte swizzle row
constexpr int TC_CLUSTER = 2;       // a CTA pair (cta_group::2): one 256-frame x 256-row UMMA spans both SMs; each CTA stages its
                                    // own 128 frames and HALF of the basis tile, the tensor cores read the other half from the peer.
                                    // (Measured: clusters of 2 or 4 pairs multicasting the basis tile cut the L2 reads by 25 / 37 %,
                                    // but only 132 / 120 SMs can hold whole clusters -- 2.35 / 2.53 ms against 2.36 ms for one pair.)
constexpr int TC_W_BYTES = TC_BM * TC_BK * 4;                    // 32 KB: the hi (or lo) image of the whole basis tile
constexpr int TC_WH_BYTES = TC_W_BYTES / TC_CLUSTER;             // 16 KB: this CTA's rows of it
constexpr int TC_X_BYTES = TC_BN * TC_BK * 4;                    // 16 KB: hi (or lo) image of the CTA's frames tile
// Shared memory: a ring of 32 KB slots -- one K-block of the basis (this CTA's rows, hi | lo) or of the frames (hi | lo)
// -- plus the RESIDENT frames operand of the scale part: its K = 85 + 1 is three K-blocks = 96 KB, loaded once per
// 128-frame tile and reused by all 61 scale tiles, so a scale K-block streams 32 KB instead of 64 (the kernel is bound
// by operand traffic from L2 plus its own 94 KB/frame of stores).  Rotation K-blocks (K = 180 + 1: six blocks, 192 KB
// of frames) stream both operands.
constexpr int TC_SLOT_BYTES = 2 * TC_WH_BYTES;                   // = 2 * TC_X_BYTES = 32 KB
constexpr int TC_SLOTS = 4;
constexpr int TC_XS_KB = 3;                                      // resident K-blocks of the scale part's frames operand
static_assert(2 * TC_X_BYTES == TC_SLOT_BYTES, "one slot holds either operand's K-block");
constexpr int TC_EPI_WARPS = 8;      // two warps per TMEM lane quarter, each drains half of the columns
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);   // warp 0 TMA producer, warp 1 MMA issuer, then the epilogue warps
constexpr int TC_TMEM_COLS = 512;   // two 256-column fp32 accumulators

// float index of element (row r, k) inside a [rows x 32] K-major SWIZZLE_128B tile image:
// 8-row groups of 1024 bytes, 16-byte chunk index XORed with the row index inside the group
__host__ __device__ inline int swz(int r, int k) {
    return (r >> 3) * 256 + (r & 7) * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
}
__host__ __device__ inline float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#else
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    float h;
    std::memcpy(&h, &u, 4);
    return h;
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // arrivals come from the peer CTA
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// remote arrive, default (CTA-scope release) semantics as in CUTLASS' ClusterBarrier::arrive(cta_id): the thread has no
// memory operations of its own to publish, and a cluster-scope release costs a MEMBAR.GPU per stage
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the same without cluster-scope release: the accumulator hand-back orders tensor-memory reads (tcgen05.wait::ld has
// completed them), not the thread's global stores -- a release fence here would wait for all 128 of them
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 >> 4 between 8-row groups,
// version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// D[tmem of both CTAs] (+)= A[smem] * B[smem]^T, kind::tf32, M = 256 (128 per CTA), N = 256 (128 rows of B per CTA), K = 8;
// issued by the pair's leader only, the descriptors name the same shared-memory offsets in both CTAs
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// Both GEMMs (part 0 = scale basis, part 1 = rotation basis) run in ONE launch with their tiles interleaved
// scale, scale, rotation: a scale tile (K = 96) is bound by its 128 KB of stores, a rotation tile (K = 192, same
// stores, twice the MMAs) by the tensor pipe, so next to each other they overlap.
struct GemmParams {
    const float *w_img[2];   // [m_tiles][kb][hi,lo][256x32 swizzled]
    const float *x_img[2];   // [n_tiles][kb][hi,lo][128x32 swizzled]
    float *out;              // [tiles of COMPACT_TILE frames][out_stride][COMPACT_TILE]
    long long out_stride;    // slots per frame (scale part + rotation part)
    int part_off[2];         // first slot of the part: GEMM row r is slot part_off + r (the bias rides in the GEMM)
    int m_tiles[2], kb[2];
    int n_frames, n_tiles;   // n_tiles = frame tiles, a multiple of TC_CLUSTER (the images are padded)
};

struct TileInfo { int part, m, n; };
// unit t of the walk: frame-tile group n = t / (m_tiles[0] + m_tiles[1]); inside it scale, scale, rotation, ... while both last
__device__ __forceinline__ TileInfo tile_info(const GemmParams &P, int t) {
    const int per_n = P.m_tiles[0] + P.m_tiles[1];
    TileInfo ti;
    ti.n = t / per_n;
    const int q = t - ti.n * per_n;
    const int triples = min(P.m_tiles[0] / 2, P.m_tiles[1]);
    if (q < 3 * triples) {
        const int tr = q / 3, r = q - 3 * tr;
        ti.part = r == 2;
        ti.m = r == 2 ? tr : 2 * tr + r;
    } else {                                   // leftovers of the longer list
        const int rest = q - 3 * triples, left0 = P.m_tiles[0] - 2 * triples;
        ti.part = rest >= left0;
        ti.m = ti.part ? triples + rest - left0 : 2 * triples + rest;
    }
    return ti;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, N = 256 rows, M = 256 frames (the pair)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 3) << 17) |
                              ((uint32_t)((TC_BN * TC_CLUSTER) >> 4) << 24);

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1) k_decode_tc(GemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t crank = blockIdx.x % TC_CLUSTER, cid = blockIdx.x / TC_CLUSTER, n_clusters = gridDim.x / TC_CLUSTER;
    constexpr uint16_t CMASK = (uint16_t)((1u << TC_CLUSTER) - 1u);
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slots = smem, *xs = smem + TC_SLOTS * TC_SLOT_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(xs + TC_XS_KB * TC_SLOT_BYTES);
    // per slot -- full: this CTA's copy has landed; peer (leader only): the follower's has; empty: the pair's MMAs have read it.
    // The same three for the resident operand (index TC_SLOTS), then the two accumulators' full / empty.
    constexpr int NB = TC_SLOTS + 1;
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NB), bar_peer = smem_u32(bars + 2 * NB);
    const uint32_t bar_tfull = smem_u32(bars + 3 * NB), bar_tempty = smem_u32(bars + 3 * NB + 2);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * NB + 4);
    const uint32_t xs_full = bar_full + 8 * TC_SLOTS, xs_empty = bar_empty + 8 * TC_SLOTS, xs_peer = bar_peer + 8 * TC_SLOTS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool resident = P.kb[0] <= TC_XS_KB;       // else the scale part streams its frames operand like the rotation part

    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_peer + 8 * s, 1); }
        // the leader's accumulator is free once the epilogue warps of BOTH CTAs have drained their halves
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, TC_CLUSTER * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // the same warp of both CTAs allocates the pair's tensor memory and later frees it
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync();                                // every CTA's barriers exist before a peer signals them
    const uint32_t tmem_base = *tmem_slot;
    // a pair walks a CONTIGUOUS range of units (frame-tile pair n major, the basis tiles inside it scale, scale, rotation, ...)
    // so that the resident frames operand changes once per m_tiles[0] + m_tiles[1] units
    const long long n_units = (long long)(P.m_tiles[0] + P.m_tiles[1]) * (P.n_tiles / TC_CLUSTER);
    const int t_begin = (int)(n_units * cid / n_clusters), t_end = (int)(n_units * (cid + 1) / n_clusters);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
        if (lane == 0) {
            uint32_t it = 0, xit = 0;
            int cur_n = -1;
            auto slot_begin = [&]() -> uint32_t {
                const uint32_t s = it % TC_SLOTS, ph = (it / TC_SLOTS) & 1u;
                ++it;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);                // the pair's MMAs are done with the slot
                mbar_arrive_expect_tx(bar_full + 8 * s, TC_SLOT_BYTES);
                return s;
            };
            for (int tile = t_begin; tile < t_end; ++tile) {
                const TileInfo ti = tile_info(P, tile);
                const int kbs = P.kb[ti.part];
                const uint8_t *w = reinterpret_cast<const uint8_t *>(P.w_img[ti.part] + (size_t)ti.m * kbs * (2 * TC_BM * TC_BK)) + crank * TC_WH_BYTES;
                const float *x = P.x_img[ti.part] + (size_t)(ti.n * TC_CLUSTER + crank) * kbs * (2 * TC_BN * TC_BK);
                const bool stream_x = ti.part == 1 || !resident;
                if (!stream_x && ti.n != cur_n) {                     // new frame tile: reload the resident operand
                    mbar_wait(xs_empty, (xit & 1u) ^ 1u);             // every MMA that read the old one has completed
                    mbar_arrive_expect_tx(xs_full, (uint32_t)kbs * TC_SLOT_BYTES);
                    for (int kb = 0; kb < kbs; ++kb)
                        tma_bulk_g2s(smem_u32(xs + kb * TC_SLOT_BYTES), x + (size_t)kb * (2 * TC_BN * TC_BK), TC_SLOT_BYTES, xs_full);
                    ++xit;
                    cur_n = ti.n;
                }
                for (int kb = 0; kb < kbs; ++kb) {
                    const uint8_t *wk = w + (size_t)kb * (2 * TC_W_BYTES);
                    const uint32_t s = slot_begin(), dst = smem_u32(slots + s * TC_SLOT_BYTES);
                    tma_bulk_g2s(dst, wk, TC_WH_BYTES, bar_full + 8 * s);                              // rows 128 crank .. of W_hi
                    tma_bulk_g2s(dst + TC_WH_BYTES, wk + TC_W_BYTES, TC_WH_BYTES, bar_full + 8 * s);   // the same rows of W_lo
                    if (stream_x) {
                        const uint32_t s2 = slot_begin();
                        tma_bulk_g2s(smem_u32(slots + s2 * TC_SLOT_BYTES), x + (size_t)kb * (2 * TC_BN * TC_BK), TC_SLOT_BYTES, bar_full + 8 * s2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank != 0) {
            // ------------------------------------------------------------------ follower: tell the leader what has landed here
            const uint32_t peer0 = mapa(bar_peer, 0), xs_peer0 = mapa(xs_peer, 0);
            uint32_t it = 0, xit = 0;
            int cur_n = -1;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const TileInfo ti = tile_info(P, tile);
                const bool stream_x = ti.part == 1 || !resident;
                if (!stream_x && ti.n != cur_n) {
                    mbar_wait(xs_full, xit & 1u);
                    mbar_arrive_cluster(xs_peer0);
                    ++xit;
                    cur_n = ti.n;
                }
                const int n_slots = P.kb[ti.part] * (stream_x ? 2 : 1);
                for (int k = 0; k < n_slots; ++k, ++it) {
                    const uint32_t s = it % TC_SLOTS, ph = (it / TC_SLOTS) & 1u;
                    mbar_wait(bar_full + 8 * s, ph);
                    mbar_arrive_cluster(peer0 + 8 * s);
                }
            }
        } else if (lane == 0) {
            // ------------------------------------------------------------------ leader: MMA issuer of the pair (one thread)
            uint32_t it = 0, tc = 0, xit = 0;
            int cur_n = -1;
            auto slot_ready = [&]() -> uint32_t {
                const uint32_t s = it % TC_SLOTS, ph = (it / TC_SLOTS) & 1u;
                ++it;
                mbar_wait(bar_full + 8 * s, ph);
                mbar_wait_cluster(bar_peer + 8 * s, ph);
                return s;
            };
            for (int tile = t_begin; tile < t_end; ++tile, ++tc) {
                const TileInfo ti = tile_info(P, tile);
                const bool stream_x = ti.part == 1 || !resident;
                const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
                if (!stream_x && ti.n != cur_n) {
                    if (cur_n >= 0) tc_commit_multicast(xs_empty, CMASK);   // both producers: the old resident operand has been read
                    mbar_wait(xs_full, xit & 1u);
                    mbar_wait_cluster(xs_peer, xit & 1u);
                    ++xit;
                    cur_n = ti.n;
                }
                mbar_wait_cluster(bar_tempty + 8 * acc, aph ^ 1u);   // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_BM;
                const int kbs = P.kb[ti.part];
                for (int kb = 0; kb < kbs; ++kb) {
                    const uint32_t sw = slot_ready(), sx = stream_x ? slot_ready() : 0u;
                    tc_fence_after();
                    const uint32_t wbase = smem_u32(slots + sw * TC_SLOT_BYTES);
                    const uint32_t xbase = stream_x ? smem_u32(slots + sx * TC_SLOT_BYTES) : smem_u32(xs + kb * TC_SLOT_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {          // UMMA K = 8 floats = 32 bytes
                        const uint64_t w_hi = umma_desc(wbase + k * 32), w_lo = umma_desc(wbase + TC_WH_BYTES + k * 32);
                        const uint64_t x_hi = umma_desc(xbase + k * 32), x_lo = umma_desc(xbase + TC_X_BYTES + k * 32);
                        umma_tf32(d_tmem, x_hi, w_hi, TC_IDESC, (kb | k) != 0);      // A = frames (M), B = basis rows (N)
                        umma_tf32(d_tmem, x_lo, w_hi, TC_IDESC, 1u);
                        umma_tf32(d_tmem, x_hi, w_lo, TC_IDESC, 1u);
                    }
                    tc_commit_multicast(bar_empty + 8 * sw, CMASK);  // both producers: the slot has been read
                    if (stream_x) tc_commit_multicast(bar_empty + 8 * sx, CMASK);
                }
                tc_commit_multicast(bar_tfull + 8 * acc, CMASK);     // both epilogues: accumulator complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
        const int lane_grp = warp & 3;                              // TMEM lanes = frames 32*lane_grp .. +31 of the 128-frame tile
        const int col_half = (warp - 2) >> 2;                       // which 64 of the 128 columns (basis rows) it drains
        const uint32_t tempty0 = mapa(bar_tempty, 0);
        uint32_t tc = 0;
        for (int tile = t_begin; tile < t_end; ++tile, ++tc) {
            const TileInfo ti = tile_info(P, tile);
            const int m = ti.m, n = ti.n * TC_CLUSTER + (int)crank;
            const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
            // this warp's 32 frames inside the compact dgrad's tiles of COMPACT_TILE frames
            const int frame0 = n * TC_BN + lane_grp * 32;
            const bool live = frame0 < P.n_frames;                  // 32-frame groups past the batch are not stored
            float *out_tile = P.out + ((size_t)(frame0 / COMPACT_TILE) * P.out_stride + P.part_off[ti.part] + (size_t)m * TC_BM +
                                       col_half * (TC_BM / 2)) * COMPACT_TILE + frame0 % COMPACT_TILE + lane;
            mbar_wait(bar_tfull + 8 * acc, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + acc * TC_BM + col_half * (TC_BM / 2);
#pragma unroll 1
            for (int chunk = 0; chunk < TC_BM / 64; ++chunk) {
                uint32_t v[32];
                tmem_ld32(taddr + chunk * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (live) {
                    float *dst = out_tile + chunk * 32 * COMPACT_TILE;      // 32 slots further
#pragma unroll
                    for (int c = 0; c < 32; ++c) __stcs(dst + c * COMPACT_TILE, __uint_as_float(v[c]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                // no CTA leaves while the pair's MMAs may still read its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// coefficients [n_frames, K] -> hi/lo tile images [n_tiles][kb][hi,lo][128 x 32 swizzled] (zero padded)
__global__ void k_split_coeffs(const float *__restrict__ x, int K, int n_frames, int kb_count, float *__restrict__ img) {
    const long long total = (long long)gridDim.y * TC_BN * kb_count * TC_BK;     // per n-tile row block
    (void)total;
    const int n_tile = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TC_BN * kb_count * TC_BK; i += gridDim.x * blockDim.x) {
        const int r = i / (kb_count * TC_BK), kk = i - r * (kb_count * TC_BK);
        const int kb = kk / TC_BK, k = kk - kb * TC_BK;
        const int frame = n_tile * TC_BN + r, kg = kb * TC_BK + k;
        // column K is the constant 1 that multiplies the means stored as column K of the basis images
        const float v = kg == K ? 1.f : ((frame < n_frames && kg < K) ? x[(long long)frame * K + kg] : 0.f);
        const float hi = tf32_hi(v);
        float *tile = img + ((size_t)n_tile * kb_count + kb) * (2 * TC_BN * TC_BK);
        tile[swz(r, k)] = hi;
        tile[TC_BN * TC_BK + swz(r, k)] = v - hi;
    }
}

}  // namespace

int tc_kblocks(int K) { return (K + 1 + TC_BK - 1) / TC_BK; }   // + the bias column
static int tc_frame_tiles(int n_frames) { return ((n_frames + TC_BN - 1) / TC_BN + TC_CLUSTER - 1) / TC_CLUSTER * TC_CLUSTER; }
size_t tc_ximg_floats(int n_frames, int K) {
    return (size_t)tc_frame_tiles(n_frames) * tc_kblocks(K) * 2 * TC_BN * TC_BK;
}

// Host: pre-split, pre-tiled basis images; GEMM row r = compact slot r of the basis' part, rows_src[r] = row of the
// [*, K] basis W it reproduces (-1: nothing, the slot decodes to 0); the means become column K.
int tc_build_basis(const float *W, const float *mean, int K, const std::vector<int32_t> &rows_src, std::vector<float> &img) {
    const int rows = (int)rows_src.size(), m_tiles = (rows + TC_BM - 1) / TC_BM, kbs = tc_kblocks(K);
    img.assign((size_t)m_tiles * kbs * 2 * TC_BM * TC_BK, 0.f);
    for (int r = 0; r < rows; ++r) {
        const int m = r / TC_BM, rl = r % TC_BM, src = rows_src[r];
        if (src < 0) continue;
        for (int k = 0; k <= K; ++k) {
            const float v = k < K ? W[(size_t)src * K + k] : mean[src], hi = tf32_hi(v);
            float *tile = &img[((size_t)m * kbs + k / TC_BK) * (2 * TC_BM * TC_BK)];
            tile[swz(rl, k % TC_BK)] = hi;
            tile[TC_BM * TC_BK + swz(rl, k % TC_BK)] = v - hi;
        }
    }
    return m_tiles;
}
int tc_rows_per_tile() { return TC_BM; }

// How many CTA pairs can be resident at once (a cluster must fit inside one GPC); kept in the plan.
cudaError_t configure_decode_tc(DevicePlan &d) {
    const size_t smem = (size_t)(TC_SLOTS + TC_XS_KB) * TC_SLOT_BYTES + 256 + 1024;
    cudaError_t e = cudaFuncSetAttribute(k_decode_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(d.sm_count / TC_CLUSTER * TC_CLUSTER));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TC_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_decode_tc, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = d.sm_count / TC_CLUSTER; }
    d.decode_max_clusters = n;
    return cudaSuccess;
}

cudaError_t launch_decode_tc(const DevicePlan &d, const float *coeff_scale, const float *coeff_rotat, int n_frames,
                             float *ximg_scale, float *ximg_rotat, float *dgrad_out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int n_tiles = tc_frame_tiles(n_frames);                  // padded to whole clusters
    const size_t smem = (size_t)(TC_SLOTS + TC_XS_KB) * TC_SLOT_BYTES + 256 + 1024;
    {
        cudaError_t e = cudaFuncSetAttribute(k_decode_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long stride = d.compact_stride;
    for (int g = 0; g < 2; ++g) {
        const float *x = g == 0 ? coeff_scale : coeff_rotat;
        const int K = g == 0 ? d.k_scale : d.k_rotat;
        k_split_coeffs<<<dim3(8, (unsigned)n_tiles), 256, 0, stream>>>(x, K, n_frames, tc_kblocks(K), g == 0 ? ximg_scale : ximg_rotat);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    GemmParams P{{d.tc_w_scale, d.tc_w_rotat}, {ximg_scale, ximg_rotat}, dgrad_out, stride, {0, d.compact_s_rows},
                 {d.tc_mt_scale, d.tc_mt_rotat}, {tc_kblocks(d.k_scale), tc_kblocks(d.k_rotat)}, n_frames, n_tiles};
    int grid = (P.m_tiles[0] + P.m_tiles[1]) * n_tiles;
    // persistent: as many clusters as can be resident at once (configure_decode_tc)
    const int max_clusters = d.decode_max_clusters > 0 ? d.decode_max_clusters : d.sm_count / TC_CLUSTER;
    if (grid > max_clusters * TC_CLUSTER) grid = max_clusters * TC_CLUSTER;
    grid = grid / TC_CLUSTER * TC_CLUSTER;
    k_decode_tc<<<grid, TC_THREADS, smem, stream>>>(P);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace sdfa
