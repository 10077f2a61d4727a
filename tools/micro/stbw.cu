// Store-bandwidth probe: how fast can a kernel stream writes with 4-byte vs 16-byte stores per thread,
// and with few warps per SM (like a GEMM epilogue) vs full occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int V> __global__ void k(float *p, size_t n_per_block, int iters_unused) {
    float *base = p + (size_t)blockIdx.x * n_per_block;
    if (V == 1) { for (size_t i = threadIdx.x; i < n_per_block; i += blockDim.x) __stcs(base + i, 1.f); }
    else { float4 v = make_float4(1, 1, 1, 1); float4 *b4 = reinterpret_cast<float4 *>(base);
           for (size_t i = threadIdx.x; i < n_per_block / 4; i += blockDim.x) __stcs(b4 + i, v); }
}
// lane = "frame", each thread writes 32 values 128 B apart (the K1 epilogue pattern): warp instr = one 128 B line
__global__ void k_lines(float *p, size_t n_per_block) {
    float *base = p + (size_t)blockIdx.x * n_per_block + (threadIdx.x & 31);
    int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (size_t line = warp * 32; line < n_per_block / 32; line += nw * 32) {
#pragma unroll
        for (int c = 0; c < 32; ++c) __stcs(base + (line + c) * 32, 1.f);
    }
}
int main() {
    size_t n = (size_t)1 << 30; float *p; cudaMalloc(&p, n * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        launch(); cudaDeviceSynchronize(); float best = 1e9;
        for (int i = 0; i < 5; ++i) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
        printf("%-40s %.3f ms  %.0f GB/s\n", name, best, n * 4 / best / 1e6);
    };
    for (int blocks : {148, 148 * 4, 148 * 16}) for (int threads : {256, 1024}) {
        char nm[64];
        snprintf(nm, 64, "st.32  blocks %d threads %d", blocks, threads); run(nm, [&] { k<1><<<blocks, threads>>>(p, n / blocks, 0); });
        snprintf(nm, 64, "st.128 blocks %d threads %d", blocks, threads); run(nm, [&] { k<4><<<blocks, threads>>>(p, n / blocks, 0); });
        snprintf(nm, 64, "lines  blocks %d threads %d", blocks, threads); run(nm, [&] { k_lines<<<blocks, threads>>>(p, n / blocks); });
    }
    return 0;
}
