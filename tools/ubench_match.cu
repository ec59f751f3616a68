// Developer tool: cost of MATCH.ANY as a function of the equality structure of the 32 lane values.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_match tools/ubench_match.cu && ./ubench_match
// Every run: 148 x 8 CTAs of 128 threads, 2048 trips of 8 independent match.any; the lane values are
// pattern[lane] ^ c with a trip-dependent c (same equality structure every trip).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(const uint32_t *__restrict__ pattern, int iters, uint32_t *sink, int use_match)
{
    const int lane = threadIdx.x & 31;
    const uint32_t base = pattern[lane];
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t x = base ^ (uint32_t)((it * 8 + u) & 0xFF);
            acc += use_match ? __match_any_sync(0xFFFFFFFFu, x) : x * 3;
        }
    }
    if (acc == 0x12345) sink[0] = acc;
}

static double run(const uint32_t *h, uint32_t *d_pat, uint32_t *sink)
{
    cudaMemcpy(d_pat, h, 32 * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<148 * 8, 128>>>(d_pat, 64, sink, 1);
    cudaEventRecord(e0);
    k<<<148 * 8, 128>>>(d_pat, 2048, sink, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3 * 1.965e9 / (8.0 * 4 * 2048 * 8);
}

int main()
{
    uint32_t *d_pat, *sink, h[32];
    cudaMalloc(&d_pat, 128);
    cudaMalloc(&sink, 4);
    auto show = [&](const char *name) { printf("%-34s %7.2f cycles per match per SM\n", name, run(h, d_pat, sink)); };
    for (int l = 0; l < 32; l++) h[l] = l;
    show("32 distinct");
    for (int l = 0; l < 32; l++) h[l] = 7;
    show("all same");
    for (int l = 0; l < 32; l++) h[l] = l >> 1;
    show("16 adjacent pairs");
    for (int l = 0; l < 32; l++) h[l] = l & 15;
    show("16 pairs at lane distance 16");
    for (int l = 0; l < 32; l++) h[l] = l >> 2;
    show("8 adjacent quads");
    for (int l = 0; l < 32; l++) h[l] = l & 7;
    show("8 strided quads");
    for (int l = 0; l < 32; l++) h[l] = l >> 4;
    show("2 halves");
    for (int l = 0; l < 32; l++) h[l] = l; h[17] = 3;
    show("1 pair + 30 singles");
    for (int l = 0; l < 32; l++) h[l] = l; h[17] = 3; h[20] = 5; h[30] = 9; h[31] = 11;
    show("4 pairs + 24 singles");
    for (int l = 0; l < 32; l++) h[l] = l; for (int l = 16; l < 24; l++) h[l] = l - 16;
    show("8 pairs + 16 singles");
    for (int l = 0; l < 32; l++) h[l] = l; h[9] = 3; h[21] = 3;
    show("1 triple + 29 singles");
    for (int l = 0; l < 32; l++) h[l] = l < 16 ? (l & 3) : l;
    show("4 quads + 16 singles");
    for (int l = 0; l < 32; l++) h[l] = l * 0x01010101u;
    show("32 distinct, 32-bit values");
    for (int l = 0; l < 32; l++) h[l] = (l >> 1) * 0x01010101u;
    show("16 pairs, 32-bit values");
    srand(7);
    for (int t = 0; t < 4; t++) {
        uint8_t b[96];
        int pos = 0;
        while (pos < 96) {
            if (rand() & 1) { int n = 1 + rand() % 8; while (n-- && pos < 96) b[pos++] = (uint8_t)rand(); }
            else if (pos) { int n = 3 + rand() % 10, d = 1 + rand() % (pos < 32 ? pos : 32); while (n-- && pos < 96) { b[pos] = b[pos - d]; ++pos; } }
        }
        int groups = 0, multi = 0, cnt[256] = {0};
        for (int l = 0; l < 32; l++) { h[l] = b[64 + l]; cnt[b[64 + l]]++; }
        for (int v = 0; v < 256; v++) { groups += cnt[v] > 0; multi += cnt[v] > 1; }
        char name[64];
        snprintf(name, sizeof name, "workload chunk (%d groups, %d multi)", groups, multi);
        show(name);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
