// Developer tool: which execution pipe the warp-collective instructions of the compress kernel use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu
//   ncu --metrics sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,\
//       sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,\
//       sm__inst_executed_pipe_uniform.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max ./ubench_pipes
// Every kernel runs 148 x 8 CTAs of 128 threads, 4096 trips of 8 independent ops of one kind.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { MATCH = 0, BALLOT, SHFL, REDUX, SYNCWARP, LDSB, BRANCHY, MATCH_RANDOM, ACTIVEMASK, NONE, MATCH_PAIRS, MATCH_MIXED, MATCH_WORKLOAD };

template <int OP>
__global__ void __launch_bounds__(128) k(int iters, uint32_t *sink, const uint32_t *__restrict__ seed)
{
    __shared__ uint32_t tab[4][256];
    __shared__ uint8_t wl[64][32];      // workload-like bytes: literal runs of 1..8 random bytes, copies of 3..12 from distance 1..32
    if (threadIdx.x == 0) {
        uint64_t s = 0x9E3779B97F4A7C15ull + blockIdx.x;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); };
        uint8_t *b = &wl[0][0];
        int pos = 0;
        while (pos < 2048) {
            if (rnd() & 1) { int n = 1 + rnd() % 8; while (n-- && pos < 2048) b[pos++] = (uint8_t)rnd(); }
            else { int n = 3 + rnd() % 10, d = 1 + rnd() % (pos < 32 ? (pos ? pos : 1) : 32); while (n-- && pos < 2048 && pos) { b[pos] = b[pos - d]; ++pos; } if (!pos) b[pos++] = (uint8_t)rnd(); }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 256; i += 32) tab[warp][i] = i * 2654435761u;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = (lane * 2654435761u >> (24 - u)) + seed[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t x = a[u] & 255u, r = 0;
            if (OP == MATCH) r = __match_any_sync(0xFFFFFFFFu, x & 7u);
            else if (OP == MATCH_RANDOM) r = __match_any_sync(0xFFFFFFFFu, x);
            else if (OP == MATCH_PAIRS) r = __match_any_sync(0xFFFFFFFFu, (x >> 1) + (lane >> 1) * 2654435761u);
            else if (OP == MATCH_MIXED) r = __match_any_sync(0xFFFFFFFFu, lane < 16 ? (x & 3u) : x + lane * 977u);
            else if (OP == MATCH_WORKLOAD) r = __match_any_sync(0xFFFFFFFFu, wl[(it * 8 + u) & 63][lane]);
            else if (OP == BALLOT) r = __ballot_sync(0xFFFFFFFFu, x & 1);
            else if (OP == SHFL) r = __shfl_sync(0xFFFFFFFFu, x, (lane + 1) & 31);
            else if (OP == REDUX) r = __reduce_add_sync(0xFFFFFFFFu, x);
            else if (OP == SYNCWARP) { tab[warp][(lane + u) & 255] = x; __syncwarp(); r = tab[warp][(lane + u + 1) & 255]; }
            else if (OP == LDSB) r = tab[warp][x];
            else if (OP == BRANCHY) { if (x & 1) r = x * 3 + 1; else r = x >> 1; if (r & 2) r ^= a[(u + 1) & 7]; }
            else if (OP == ACTIVEMASK) r = __activemask();
            else r = x * 5;
            a[u] += r + it;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += a[u];
    if (acc == 0x12345) sink[0] = acc;
}

template <int OP>
void run(const char *name, uint32_t *sink, uint32_t *seed)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<148 * 8, 128>>>(64, sink, seed);
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 128>>>(4096, sink, seed);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // ops per SM per cycle at 1.965 GHz: 8 CTAs x 4 warps x 4096 x 8 ops per SM
    const double ops = 8.0 * 4 * 4096 * 8;
    printf("%-14s %8.3f ms  %6.2f cycles per op per SM (at 1965 MHz)\n", name, ms, ms * 1e-3 * 1.965e9 / ops);
}

int main()
{
    uint32_t *sink, *seed;
    cudaMalloc(&sink, 4);
    cudaMalloc(&seed, 4);
    cudaMemset(seed, 0, 4);
    run<NONE>("none", sink, seed);
    run<MATCH>("match8", sink, seed);
    run<MATCH_RANDOM>("match_random", sink, seed);
    run<MATCH_PAIRS>("match_pairs", sink, seed);
    run<MATCH_MIXED>("match_mixed", sink, seed);
    run<MATCH_WORKLOAD>("match_workload", sink, seed);
    run<BALLOT>("ballot", sink, seed);
    run<SHFL>("shfl", sink, seed);
    run<REDUX>("redux", sink, seed);
    run<SYNCWARP>("sts_sync_lds", sink, seed);
    run<LDSB>("lds", sink, seed);
    run<BRANCHY>("branchy", sink, seed);
    run<ACTIVEMASK>("activemask", sink, seed);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
