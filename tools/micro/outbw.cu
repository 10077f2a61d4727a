// Store-pattern probe for the output kernel: CTA = (64 vertices = 192 floats, 32 frames), one 128-byte store per warp and
// frame, rows of R floats.  R = 15069 (FLAME: rows start at any 4-byte offset inside a line) against R = 15072 / 15104
// (rows start on a sector / line boundary): is the 60 % of the HBM peak a matter of store alignment?
#include <algorithm>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192) k(float *out, long long R, int n_frames, int shift_mode) {
    const int chunk = blockIdx.x, frame0 = blockIdx.y * 32, e = threadIdx.x;
    for (int f = 0; f < 32 && frame0 + f < n_frames; ++f) {
        long long base = (long long)(frame0 + f) * R;
        long long g = (long long)chunk * 192 + e;
        if (shift_mode) g -= base % 32;                       // windows shifted so that every warp store is line aligned
        if (g >= 0 && g < R) __stcs(out + base + g, 1.f);
    }
}
// CTA = (W floats of every row, FG frames): thread = element(s), one 128-byte store per warp, frame and 32 elements
__global__ void k_piece(float *out, long long R, int n_frames, int W, int FG) {
    const int frame0 = blockIdx.y * FG;
    for (int e = threadIdx.x; e < W; e += blockDim.x) {
        const long long g = (long long)blockIdx.x * W + e;
        if (g >= R) break;
        for (int f = 0; f < FG && frame0 + f < n_frames; ++f) __stcs(out + (long long)(frame0 + f) * R + g, 1.f);
    }
}
// CTA = FB whole frames: one contiguous block of FB * R floats, written front to back
__global__ void __launch_bounds__(1024) k_rows(float *out, long long R, int n_frames, int FB) {
    const long long base = (long long)blockIdx.x * FB * R, total = (long long)min(FB, n_frames - blockIdx.x * FB) * R;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) __stcs(out + base + i, 1.f);
}
int main() {
    const int n = 75600;
    float *out; cudaMalloc(&out, (size_t)n * 15104 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 2; ++mode)
        for (long long R : {15069LL, 15072LL, 15104LL}) {
            dim3 grid((unsigned)((R + 191) / 192 + mode), (n + 31) / 32);
            float best = 1e9;
            for (int i = 0; i < 4; ++i) { cudaEventRecord(a); k<<<grid, 192>>>(out, R, n, mode); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
            printf("R = %lld floats, %s windows: %.3f ms  %.0f GB/s  %s\n", R, mode ? "line-aligned (shifted)" : "fixed", best, (double)n * R * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int W : {192, 384, 768, 1536, 3072, 15069})
        for (int FG : {8, 32}) {
            const int threads = W >= 768 ? 768 : W;
            dim3 grid((unsigned)((15069 + W - 1) / W), (unsigned)((n + FG - 1) / FG));
            float best = 1e9;
            for (int i = 0; i < 4; ++i) { cudaEventRecord(a); k_piece<<<grid, threads>>>(out, 15069, n, W, FG); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
            printf("pieces of %d floats x %d frames, %d threads: %.3f ms  %.0f GB/s  %s\n", W, FG, threads, best, (double)n * 15069 * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int FB : {1, 4, 8, 16})
        for (int threads : {256, 1024}) {
            float best = 1e9;
            for (int i = 0; i < 4; ++i) { cudaEventRecord(a); k_rows<<<(n + FB - 1) / FB, threads>>>(out, 15069, n, FB); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
            printf("CTA = %d whole frames (contiguous), %d threads: %.3f ms  %.0f GB/s  %s\n", FB, threads, best, (double)n * 15069 * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
