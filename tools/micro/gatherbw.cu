// Gather-bandwidth probe for the reference-layout dgrad (one 359 KB row per frame, 36-byte records of the active
// triangles): what the memory system delivers when (a) a CTA owns a row block (277 records scattered over the whole row)
// and 64 rows, like k_assemble_gather / k_assemble_gather2, against (b) a CTA sweeping one row front to back over ALL
// active records in address order.  Input: tools/micro/_bin/flame_blocks.bin (tools/micro/gatherbw_data.py).
#include <algorithm>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
constexpr int ROW = 89784;
__global__ void k_block(const float *d, const int *tri, const int *ptr, int n_rows, float *sink) {
    // CTA = (block, tile of 64 rows); warp w walks records w, w+16, ...: 64 rows x 9 words each, 18 loads per lane
    const int b = blockIdx.x, row0 = blockIdx.y * 64, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int r = ptr[b] + warp; r < ptr[b + 1]; r += blockDim.x >> 5) {
        const float *src = d + (size_t)row0 * ROW + (size_t)tri[r] * 9;
#pragma unroll
        for (int k = 0; k < 18; ++k) { const int i = lane + 32 * k, f = i / 9; acc += __ldg(src + (size_t)f * ROW + (i - 9 * f)); }
    }
    if (acc == 123.456f) sink[0] = acc;
}
__global__ void k_row(const float *d, const int *word, int n_words, int n_rows, int rows_per_cta, float *sink) {
    // CTA sweeps rows_per_cta rows front to back; word[] = sorted word offsets of all active records
    float acc = 0.f;
    for (int q = 0; q < rows_per_cta; ++q) {
        const int row = blockIdx.x * rows_per_cta + q;
        if (row >= n_rows) break;
        const float *src = d + (size_t)row * ROW;
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) acc += __ldg(src + word[i]);
    }
    if (acc == 123.456f) sink[0] = acc;
}
int main() {
    FILE *fp = fopen("tools/micro/_bin/flame_blocks.bin", "rb");
    if (!fp) { printf("run tools/micro/gatherbw_data.py first\n"); return 1; }
    int nb; fread(&nb, 4, 1, fp);
    std::vector<int> tri, ptr{0};
    for (int b = 0; b < nb; ++b) { int c; fread(&c, 4, 1, fp); size_t o = tri.size(); tri.resize(o + c); fread(tri.data() + o, 4, c, fp); ptr.push_back((int)tri.size()); }
    std::vector<int> uniq(tri); std::sort(uniq.begin(), uniq.end()); uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<int> word; for (int t : uniq) for (int j = 0; j < 9; ++j) word.push_back(t * 9 + j);
    const int n_rows = 16384;                                     // 5.9 GB
    float *d, *sink; int *dtri, *dptr, *dword;
    cudaMalloc(&d, (size_t)n_rows * ROW * 4); cudaMemset(d, 0, (size_t)n_rows * ROW * 4); cudaMalloc(&sink, 4);
    cudaMalloc(&dtri, tri.size() * 4); cudaMalloc(&dptr, ptr.size() * 4); cudaMalloc(&dword, word.size() * 4);
    cudaMemcpy(dtri, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dword, word.data(), word.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, double useful_bytes, auto launch) {
        launch(); cudaDeviceSynchronize(); float best = 1e9;
        for (int i = 0; i < 4; ++i) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
        printf("%-58s %.3f ms  useful %.0f GB/s  (%.2f us per row)  %s\n", name, best, useful_bytes / best / 1e6, best * 1e3 / n_rows, cudaGetErrorString(cudaGetLastError()));
    };
    printf("blocks %d, records incl. duplicates %zu, distinct active triangles %zu\n", nb, tri.size(), uniq.size());
    for (int threads : {512, 1024})
        run(threads == 512 ? "block x 64 rows (gen 1/2 pattern), 512 threads" : "block x 64 rows (gen 1/2 pattern), 1024 threads",
            (double)tri.size() * 36 * n_rows, [&] { k_block<<<dim3(nb, n_rows / 64), threads>>>(d, dtri, dptr, n_rows, sink); });
    for (int threads : {256, 512, 1024}) for (int rpc : {1, 4}) {
        char nm[96]; snprintf(nm, 96, "row sweep, all active records in order, %d thr, %d rows/CTA", threads, rpc);
        run(nm, (double)uniq.size() * 36 * n_rows, [&] { k_row<<<(n_rows + rpc - 1) / rpc, threads>>>(d, dword, (int)word.size(), n_rows, rpc, sink); });
    }
    return 0;
}
