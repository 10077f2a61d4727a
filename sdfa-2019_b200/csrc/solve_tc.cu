// solve_tc.cu -- K3T: the batched solve (A^T A + reg I) X = RHS on the 5th-generation tensor cores.
//
// Replaces `solver_.solve(...)` of the reference's per-frame path (deformation/cpp/src/deform_triangle_impl.hpp:286-292,
// Eigen::SparseLU) for templates whose plan fits (tplan.hpp): nested-dissection block Cholesky whose blocks are
// explicit small dense matrices, executed as tcgen05.mma kind::tf32 products
//      D[128 columns, N] (+)= A[128 columns, K] * Bt[N, K]
// with the right-hand-side columns (frame x coordinate) on the M dimension = the 128 lanes of tensor memory:
// A (the source rows, split into TF32 hi / lo) and D (accumulators) both live in tensor memory, so a result
// is turned into the next product's operand by one tcgen05.ld / split / tcgen05.st round trip of the epilogue
// warps and never goes through shared memory; Bt (the fixed matrices, pre-split hi / lo, K-major
// SWIZZLE_128B tile images) streams L2 -> shared-memory ring with cp.async.bulk.  3xTF32: every K-step issues
// A_hi B_hi + A_lo B_hi + A_hi B_lo into the fp32 accumulator.
//
// One persistent CTA per SM loops over 128-column tiles of scratch[tile][row][128]; warp roles:
//   warp 16   streamer  : matrix chunks -> ring (mbarrier full / empty)
//   warp 17   issuer    : one thread walks the MmaOp stream
//   warp 18   row loader: the scratch rows the EpiOp streams will add, TMA bulk copies into a second ring, ops ahead
//   warp 19   row storer: bulk-stores the rows an op wrote into its ring stage, then frees the stage
//   warps 0-7 EPI stream 0, warps 8-15 EPI stream 1: thread = column = tensor-memory lane (two warps per lane
//             quarter split an op's chunks); each walks its own ops of the EpiOp list.
// The three instruction streams synchronise through single-use-per-tile mbarrier events chosen by the planner (an op
// that only reads columns another stream will overwrite signals a separate event right after its tcgen05.ld); a named
// barrier over the EPI warps closes a tile.  The issuer works from MmaIssue records resolved once per CTA.  Per-op
// latencies (a tensor-memory round trip, the hand-off to the other stream, the issuer's per-op work) bound a tile, not
// bandwidth: two EPI streams let the round trips of independent tree nodes overlap.
#include "device_plan.hpp"

namespace sdfa {

namespace {

#ifndef TS_RING_N
#define TS_RING_N 3
#endif
#ifndef TS_STAGED_STORES
#define TS_STAGED_STORES 1       // 1: rows leave through the op's row-ring stage and a bulk store (storer warp); 0: plain stores
#endif
#ifndef TS_ROWRING_N
#define TS_ROWRING_N 3
#endif
constexpr int TS_RING = TS_RING_N;             // matrix ring stages (32 KB each)
constexpr int TS_ROWRING = TS_ROWRING_N;       // scratch-row ring stages (<= 64 rows x 128 columns each)
constexpr int TS_ROWSTAGE_BYTES = TS_MAX_NODE * TS_COLS * 4;
constexpr int TS_EPI_WARPS = 8;                // per EPI stream; two warps per tensor-memory lane quarter: the first four 8-column chunks of an op, and the rest
constexpr int TS_EPI_STREAMS = 2;
constexpr int TS_EPI_ALL = TS_EPI_STREAMS * TS_EPI_WARPS;           // warps 0 .. 15: the EPI streams (a warp's tensor-memory lane quarter is warp % 4)
constexpr int TS_W_STREAMER = TS_EPI_ALL, TS_W_ISSUER = TS_EPI_ALL + 1, TS_W_LOADER = TS_EPI_ALL + 2, TS_W_STORER = TS_EPI_ALL + 3;
constexpr int TS_THREADS = 32 * (TS_EPI_ALL + 3 + TS_STAGED_STORES);

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, K = 8 (A: lane = row, 8 consecutive 32-bit columns)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
// k (1 .. 4, warp uniform) consecutive 8-column chunks in as few tensor-memory instructions as possible: an epilogue op is
// a chain of tensor-memory round trips, and four x8 accesses cost about four times one x32 access
__device__ __forceinline__ void tmem_ld_chunks(uint32_t taddr, int k, uint32_t (&v)[32]) {
    switch (k) {
        case 4: tmem_ld32(taddr, v); break;
        case 3: tmem_ld16(taddr, v); tmem_ld8(taddr + 16, v + 16); break;
        case 2: tmem_ld16(taddr, v); break;
        case 1: tmem_ld8(taddr, v); break;
        default: break;
    }
}
__device__ __forceinline__ void tmem_zero_chunks(uint32_t taddr, int k) {
    for (int c = 0; c < k; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + 8 * c), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_chunks(uint32_t taddr, int k, const uint32_t (&v)[32]) {
    switch (k) {
        case 4: tmem_st32(taddr, v); break;
        case 3: tmem_st16(taddr, v); tmem_st8(taddr + 16, v + 16); break;
        case 2: tmem_st16(taddr, v); break;
        case 1: tmem_st8(taddr, v); break;
        default: break;
    }
}

// a profiling stamp, or (regular build) a point the compiler does not move memory operations across: keeps an op's
// phases -- and the registers they need -- apart
#ifndef TS_PROF_LIGHT
#define TS_PROF_LIGHT 0           // 1: the profiling build only stamps the start of every MMA op (least perturbation)
#endif
#define TS_STAMP(i) do { if (prof && !TS_PROF_LIGHT) P.prof[i] = clock64(); else asm volatile("" ::: "memory"); } while (0)

// An MmaOp as the issuing thread wants it: resolved shared-memory barrier addresses, absolute tensor-memory addresses,
// descriptor words relative to the ring stage -- built once per CTA (thread = op) so that the per-op path of the one
// thread every product of the tile goes through is a few loads and adds
struct MmaIssue {              // 48 bytes
    uint32_t wait0, wait1;     // barrier addresses (0: none)
    uint32_t commit;           // barrier address (0: none)
    uint32_t flags_k8;         // MmaFlags | k8 << 8
    uint32_t idesc, d, a_hi, a_lo;
    uint32_t b_hi_rel, b_lo_rel;   // (byte offset in the stage) >> 4
    uint32_t kb_step, pad;
};

struct TsParams {
    const MmaOp *mma;
    const EpiOp *epi;
    const uint8_t *matrix;
    const uint32_t *chunk_off;
    int n_mma, n_epi, n_chunks, n_mma_events, n_epi_events, n_ring_ops;
    float *scratch;            // [n_tiles][n_rows][128]
    int n_rows, n_tiles;
    long long *prof;           // optional timeline of CTA 0's second tile: 6 clocks per EPI op, then 5 per MMA op
};

// instruction descriptor: D fp32, A / B tf32, both K-major, M = 128; N is filled in per op
constexpr uint32_t TS_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
constexpr int TS_MAX_CHUNKS8 = TS_MAX_NODE / 8;
constexpr int TS_N_BARS = 2 * TS_RING + 3 * TS_ROWRING + 1 + 2 * TS_MAX_EVENTS;

template <bool PROF>
__global__ void __launch_bounds__(TS_THREADS, 1) k_solve_tc(TsParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B images: 1024-byte aligned
    uint8_t *ring = smem;
    float *rowring = reinterpret_cast<float *>(smem + TS_RING * TS_STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TS_RING * TS_STAGE_BYTES + TS_ROWRING * TS_ROWSTAGE_BYTES);
    uint64_t *bar_full = bars, *bar_empty = bars + TS_RING, *bar_rfull = bars + 2 * TS_RING, *bar_rempty = bar_rfull + TS_ROWRING;
    uint64_t *bar_sdone = bar_rempty + TS_ROWRING;
    uint64_t *bar_fwd = bar_sdone + TS_ROWRING, *bar_mma = bar_fwd + 1, *bar_epi = bar_mma + TS_MAX_EVENTS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_epi + TS_MAX_EVENTS);
    MmaIssue *mma_sm = reinterpret_cast<MmaIssue *>(tmem_slot + 4);
    EpiOp *epi_sm = reinterpret_cast<EpiOp *>(mma_sm + P.n_mma);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < P.n_epi * 8; i += TS_THREADS)
        reinterpret_cast<uint32_t *>(epi_sm)[i] = reinterpret_cast<const uint32_t *>(P.epi)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_RING; ++s) { mbar_init(smem_u32(bar_full + s), 1); mbar_init(smem_u32(bar_empty + s), 1); }
        for (int s = 0; s < TS_ROWRING; ++s) {
            mbar_init(smem_u32(bar_rfull + s), 1);
            mbar_init(smem_u32(bar_rempty + s), TS_EPI_WARPS);   // the warps of the stream whose op owns the stage, or the storer on their behalf
            mbar_init(smem_u32(bar_sdone + s), TS_EPI_WARPS);
        }
        mbar_init(smem_u32(bar_fwd), TS_STAGED_STORES ? 1 : TS_EPI_WARPS);   // whoever stores the forward sweep's rows
        for (int e = 0; e < P.n_mma_events; ++e) mbar_init(smem_u32(bar_mma + e), 1);     // tcgen05.commit
        for (int e = 0; e < P.n_epi_events; ++e) mbar_init(smem_u32(bar_epi + e), TS_EPI_WARPS);     // one arrive per warp of the op's stream
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == TS_W_ISSUER) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TS_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    for (int i = threadIdx.x; i < P.n_mma; i += TS_THREADS) {
        const MmaOp op = P.mma[i];
        MmaIssue q;
        q.wait0 = op.wait_epi >= 0 ? smem_u32(bar_epi + op.wait_epi) : 0u;
        q.wait1 = op.wait_epi2 >= 0 ? smem_u32(bar_epi + op.wait_epi2) : 0u;
        if (q.wait0 == 0u) { q.wait0 = q.wait1; q.wait1 = 0u; }
        q.commit = op.commit_mma >= 0 ? smem_u32(bar_mma + op.commit_mma) : 0u;
        q.flags_k8 = (uint32_t)op.flags | ((uint32_t)op.k8 << 8);
        q.idesc = TS_IDESC | ((uint32_t)(op.n >> 3) << 17);
        q.d = tmem_base + op.d_col;
        q.a_hi = tmem_base + op.a_hi_col;
        q.a_lo = tmem_base + op.a_lo_col;
        q.b_hi_rel = op.b_hi_off >> 4;
        q.b_lo_rel = op.b_lo_off >> 4;
        q.kb_step = (uint32_t)op.n * 8u - 6u;                         // (n * 128 - 96) >> 4: on to the next 32-wide K block image
        q.pad = 0u;
        mma_sm[i] = q;
    }
    __syncthreads();
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));

    if (warp == TS_W_STREAMER) {
        // ------------------------------------------------------------------ streamer
        if (leader) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x)
                for (int c = 0; c < P.n_chunks; ++c, ++it) {
                    const uint32_t s = it % TS_RING, ph = (it / TS_RING) & 1u;
                    mbar_wait(smem_u32(bar_empty + s), ph ^ 1u);
                    const uint32_t off = P.chunk_off[c], bytes = P.chunk_off[c + 1] - off;
                    mbar_arrive_expect_tx(smem_u32(bar_full + s), bytes);
                    tma_bulk_g2s(smem_u32(ring + s * TS_STAGE_BYTES), P.matrix + off, bytes, smem_u32(bar_full + s));
                }
        }
    } else if (warp == TS_W_ISSUER) {
        // ------------------------------------------------------------------ MMA issuer
        if (leader) {
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t par = tcount & 1u;
                uint32_t stage16 = 0, slot = 0;                      // ring stage address >> 4
                const bool prof = PROF && blockIdx.x == 0 && tcount == 1;
                MmaIssue nxt = mma_sm[0];
                for (int m = 0; m < P.n_mma; ++m) {
                    const MmaIssue op = nxt;
                    nxt = mma_sm[m + 1 < P.n_mma ? m + 1 : m];
                    if (prof) P.prof[6 * P.n_epi + 5 * m] = clock64();
                    if (op.wait0) {
                        mbar_wait(op.wait0, par);
                        if (op.wait1) mbar_wait(op.wait1, par);
                        tc_fence_after();
                    }
                    if (prof && !TS_PROF_LIGHT) P.prof[6 * P.n_epi + 5 * m + 1] = clock64();
                    if (op.flags_k8 & MMA_CHUNK_FIRST) {
                        slot = it % TS_RING;
                        mbar_wait(smem_u32(bar_full + slot), (it / TS_RING) & 1u);
                        tc_fence_after();
                        stage16 = smem_u32(ring + slot * TS_STAGE_BYTES) >> 4;
                    }
                    if (prof && !TS_PROF_LIGHT) P.prof[6 * P.n_epi + 5 * m + 2] = clock64();
                    // descriptor low words (address >> 4): +2 per K = 8 step inside a 32-wide K block, then on to the next block image
                    constexpr uint64_t DESC_HI = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
                    uint32_t a_hi = op.a_hi, a_lo = op.a_lo;
                    uint32_t b_hi = ((stage16 + op.b_hi_rel) & 0x3FFFu) | (1u << 16);
                    uint32_t b_lo = ((stage16 + op.b_lo_rel) & 0x3FFFu) | (1u << 16);
                    uint32_t acc = op.flags_k8 & MMA_ACCUMULATE;
                    const uint32_t k8 = op.flags_k8 >> 8;
                    for (uint32_t j = 0; j < k8; ++j) {
                        umma_tf32_ts(op.d, a_hi, DESC_HI | b_hi, op.idesc, acc);
                        umma_tf32_ts(op.d, a_lo, DESC_HI | b_hi, op.idesc, 1u);
                        umma_tf32_ts(op.d, a_hi, DESC_HI | b_lo, op.idesc, 1u);
                        acc = 1u;
                        a_hi += 8; a_lo += 8;
                        const uint32_t step = (j & 3u) == 3u ? op.kb_step : 2u;
                        b_hi += step; b_lo += step;
                    }
                    if (prof && !TS_PROF_LIGHT) P.prof[6 * P.n_epi + 5 * m + 3] = clock64();
                    if (op.flags_k8 & MMA_CHUNK_LAST) { tc_commit(smem_u32(bar_empty + slot)); ++it; }
                    if (op.commit) tc_commit(op.commit);
                    if (prof && !TS_PROF_LIGHT) P.prof[6 * P.n_epi + 5 * m + 4] = clock64();
                }
            }
        }
    } else if (warp == TS_W_LOADER) {
        // ------------------------------------------------------------------ row loader
        // every op that adds or stores scratch rows owns the next stage of the row ring (in list order, whichever stream runs it)
        if (leader) {
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tcount) {
                const float *sc = P.scratch + (size_t)tile * P.n_rows * TS_COLS;
                for (int e = 0; e < P.n_epi; ++e) {
                    const EpiOp op = epi_sm[e];
                    if (!(op.flags & (EPI_ADD_GLOBAL | EPI_STORE_GLOBAL))) continue;
                    // rows stored earlier in this tile (the forward sweep's u) are only read by the backward sweep:
                    // wait until those stores are complete
                    if (op.flags & EPI_AFTER_STORES) mbar_wait(smem_u32(bar_fwd), tcount & 1u);
                    const uint32_t s = it % TS_ROWRING, ph = (it / TS_ROWRING) & 1u;
                    mbar_wait(smem_u32(bar_rempty + s), ph ^ 1u);
                    if (op.flags & EPI_ADD_GLOBAL) {
                        const uint32_t bytes = (uint32_t)op.n_valid * TS_COLS * 4u;
                        mbar_arrive_expect_tx(smem_u32(bar_rfull + s), bytes);
                        tma_bulk_g2s(smem_u32(rowring + (size_t)s * (TS_ROWSTAGE_BYTES / 4)), sc + (size_t)op.row_in * TS_COLS, bytes,
                                     smem_u32(bar_rfull + s));
                    } else mbar_arrive(smem_u32(bar_rfull + s));
                    ++it;
                }
            }
        }
#if TS_STAGED_STORES
    } else if (warp == TS_W_STORER) {
        // ------------------------------------------------------------------ row storer
        if (leader) {
            uint32_t it = 0, sdone_par = 0;                          // bit s: parity of stage s's next "written" phase
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                float *sc = P.scratch + (size_t)tile * P.n_rows * TS_COLS;
                for (int e = 0; e < P.n_epi; ++e) {
                    const EpiOp op = epi_sm[e];
                    if (!(op.flags & (EPI_ADD_GLOBAL | EPI_STORE_GLOBAL))) continue;
                    const uint32_t s = it % TS_ROWRING;
                    ++it;
                    if (!(op.flags & EPI_STORE_GLOBAL)) continue;   // the op's stream frees such stages itself
                    mbar_wait(smem_u32(bar_sdone + s), (sdone_par >> s) & 1u);   // all warps have written their columns (and fenced)
                    sdone_par ^= 1u << s;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(sc + (size_t)op.row_out * TS_COLS), "r"(smem_u32(rowring + (size_t)s * (TS_ROWSTAGE_BYTES / 4))),
                                   "r"((uint32_t)op.n_valid * TS_COLS * 4u) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (op.flags & EPI_LAST_FWD_STORE) {
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the forward sweep's rows are in memory
                        mbar_arrive(smem_u32(bar_fwd));
                    } else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_rempty + s)), "r"((uint32_t)TS_EPI_WARPS) : "memory");
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
#endif
    } else {
        // ------------------------------------------------------------------ EPI streams: thread = column
        const int my_stream = warp / TS_EPI_WARPS;
        const int sw = warp % TS_EPI_WARPS;                    // warp inside the stream
        const int lane_grp = warp & 3;                               // tensor-memory lanes 32*lane_grp.. belong to this warp
        const int half = sw >> 2;                                    // this warp takes the chunks NC * half .. NC * half + NC - 1
        const int col = lane_grp * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        constexpr int NC = TS_MAX_CHUNKS8 / 2;
        static_assert(NC == 4, "a warp's share of an op is at most four chunks = one x32 tensor-memory access");
        const int c0 = half * NC;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t par = tcount & 1u;
            const bool prof = PROF && blockIdx.x == 0 && tcount == 1 && sw == 0 && lane == 0;
            const uint32_t ring_base = tcount * (uint32_t)P.n_ring_ops;
#pragma unroll 1
            for (int e = 0; e < P.n_epi; ++e) {
                if (epi_sm[e].stream != my_stream) continue;
                const EpiOp op = epi_sm[e];
                TS_STAMP(6 * e);
                const int nv = op.n_valid - c0 * 8;                  // valid rows among this warp's NC * 8
                const int k = min(NC, max(0, (int)op.n_chunks - c0));   // this warp's chunks of the op (warp uniform)
                uint32_t g[NC * 8];                                  // chunk c0 + c, row i of it at g[8 c + i] (float bits)
                const bool ring_op = (op.flags & (EPI_ADD_GLOBAL | EPI_STORE_GLOBAL)) != 0;
                const uint32_t rit = ring_base + op.ring_seq, rs = rit % TS_ROWRING;
                float *stage_col = rowring + (size_t)rs * (TS_ROWSTAGE_BYTES / 4) + c0 * 8 * TS_COLS + col;
                if (ring_op) mbar_wait(smem_u32(bar_rfull + rs), (rit / TS_ROWRING) & 1u);   // long done, normally: the loader runs ops ahead
                const bool rows_first = (op.flags & (EPI_ADD_GLOBAL | EPI_FROM_TMEM)) == EPI_ADD_GLOBAL;
                if (rows_first) {                                    // a leaf's rows: into registers before any event wait
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j) g[j] = j < nv ? __float_as_uint(stage_col[j * TS_COLS]) : 0u;
                }
                TS_STAMP(6 * e + 1);
                if (op.wait_epi >= 0) mbar_wait(smem_u32(bar_epi + op.wait_epi), par);
                if (op.wait_mma >= 0) mbar_wait(smem_u32(bar_mma + op.wait_mma), par);
                if (op.wait_epi >= 0 || op.wait_mma >= 0) tc_fence_after();
                TS_STAMP(6 * e + 2);
                if (op.flags & EPI_FROM_TMEM) {
                    tmem_ld_chunks(tlane + op.src_col + 8 * c0, k, g);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (op.signal_read >= 0) {                       // the source columns may be overwritten from here on
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(bar_epi + op.signal_read));
                    }
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j) g[j] = j < nv ? g[j] : 0u;   // padding rows / chunks this op does not have
                    if (op.flags & EPI_ZERO_SRC) tmem_zero_chunks(tlane + op.src_col + 8 * c0, k);
                } else if (!rows_first) {
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j) g[j] = 0u;
                }
                TS_STAMP(6 * e + 3);
                if ((op.flags & EPI_ADD_GLOBAL) && !rows_first) {    // the rows are staged in the row ring
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j)
                        if (j < nv) g[j] = __float_as_uint(__uint_as_float(g[j]) + stage_col[j * TS_COLS]);
                }
#if TS_STAGED_STORES
                if (op.flags & EPI_STORE_GLOBAL) {                   // into the op's ring stage; the storer bulk-stores it and frees the stage
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j)
                        if (j < nv) stage_col[j * TS_COLS] = __uint_as_float(g[j]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(bar_sdone + rs));
                } else if (ring_op) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(bar_rempty + rs));
                }
#else
                if (ring_op) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(bar_rempty + rs));
                }
                if (op.flags & EPI_STORE_GLOBAL) {                   // one 128-byte line per warp and row
                    float *dst = P.scratch + ((size_t)tile * P.n_rows + op.row_out + c0 * 8) * TS_COLS + col;
#pragma unroll
                    for (int j = 0; j < NC * 8; ++j)
                        if (j < nv) dst[j * TS_COLS] = __uint_as_float(g[j]);
                    if (op.flags & EPI_LAST_FWD_STORE) {
                        // the forward sweep's rows (all stored by this stream's threads) will be bulk-loaded again by the
                        // row loader: make them visible to the async proxy, then tell it
                        __threadfence();
                        asm volatile("fence.proxy.async;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(bar_fwd));
                    }
                }
#endif
                TS_STAMP(6 * e + 4);
                if (op.flags & EPI_ST_SPLIT) {                       // lo halves leave 16 columns at a time (registers), hi in place
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t hi = g[16 * h + j] & 0xFFFFE000u;
                            lo[j] = __float_as_uint(__uint_as_float(g[16 * h + j]) - __uint_as_float(hi)) & 0xFFFFE000u;
                            g[16 * h + j] = hi;
                        }
                        const int kh = k - 2 * h;                    // chunks of this half
                        if (kh >= 2) tmem_st16(tlane + op.lo_col + 8 * c0 + 16 * h, lo);
                        else if (kh == 1) tmem_st8(tlane + op.lo_col + 8 * c0 + 16 * h, lo);
                    }
                }
                if (op.flags & (EPI_ST_RAW | EPI_ST_SPLIT)) tmem_st_chunks(tlane + op.hi_col + 8 * c0, k, g);
                if (op.flags & (EPI_ST_RAW | EPI_ST_SPLIT | EPI_ZERO_SRC)) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                if (op.signal_epi >= 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(bar_epi + op.signal_epi));
                }
                TS_STAMP(6 * e + 5);
            }
            // tile boundary: both streams are done with tensor memory (one of them has waited for the last MMA)
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TS_EPI_ALL) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TS_W_ISSUER) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS_TMEM_COLS) : "memory");
    }
}

}  // namespace

size_t solve_tc_smem_bytes(int n_mma, int n_epi) {
    return 1024 + (size_t)TS_RING * TS_STAGE_BYTES + (size_t)TS_ROWRING * TS_ROWSTAGE_BYTES + (size_t)TS_N_BARS * 8 + 16 +
           (size_t)n_mma * sizeof(MmaIssue) + (size_t)n_epi * sizeof(EpiOp);
}

cudaError_t launch_solve_tc(const DevicePlan &d, float *scratch, int n_frames, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const size_t smem = solve_tc_smem_bytes(d.ts_n_mma, d.ts_n_epi);
    {   // per function and device, not per handle: set on every launch (another handle may need a different size)
        cudaError_t e = cudaFuncSetAttribute(k_solve_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_solve_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int n_tiles = 3 * ((n_frames + TS_COLS - 1) / TS_COLS);
    TsParams P{d.ts_mma, d.ts_epi, d.ts_matrix, d.ts_chunk_off, d.ts_n_mma, d.ts_n_epi, d.ts_n_chunks,
               d.ts_n_mma_events, d.ts_n_epi_events, d.ts_n_ring_ops, scratch, d.n_free, n_tiles, d.solve_prof};
    int grid = d.sm_count;
    if (grid > n_tiles) grid = n_tiles;
    if (d.solve_prof) k_solve_tc<true><<<grid, TS_THREADS, smem, stream>>>(P);     // per-op clocks of one tile (SDFA_SOLVE_PROFILE=1)
    else k_solve_tc<false><<<grid, TS_THREADS, smem, stream>>>(P);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sdfa
