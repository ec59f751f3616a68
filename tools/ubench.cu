// Micro-benchmarks of the warp primitives the compress kernel leans on (developer tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
// Prints cycles per warp-instruction at 1..16 warps per SM sub-partition, for a dependent chain
// (latency) and for 8 independent ops per trip (throughput), on three data patterns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { MATCH = 0, BALLOT = 1, SHFL = 2, ATOMS_OR = 3, REDUX = 4, LDS_RAND = 5 };

template <int OP, bool DEP>
__global__ void k(uint32_t pattern, int iters, unsigned long long *out, uint32_t *sink)
{
    __shared__ uint32_t tab[16][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 256; i += 32) tab[warp & 15][i] = i;
    __syncthreads();
    uint32_t v = pattern == 0 ? lane : pattern == 1 ? 7u : (lane * 2654435761u >> 24);
    uint32_t a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = v + u * 3;
    uint32_t acc = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t x = DEP ? (a[0] & 255u) : (a[u] & 255u);
            uint32_t r;
            if (OP == MATCH) r = __match_any_sync(0xFFFFFFFFu, x);
            else if (OP == BALLOT) r = __ballot_sync(0xFFFFFFFFu, x & 1);
            else if (OP == SHFL) r = __shfl_sync(0xFFFFFFFFu, x, (lane + 1) & 31);
            else if (OP == ATOMS_OR) r = atomicOr(&tab[warp & 15][x], 1u << lane);
            else if (OP == REDUX) r = __reduce_add_sync(0xFFFFFFFFu, x);
            else r = tab[warp & 15][x];
            if (DEP) a[0] = (r ^ a[0]) + it; else a[u] += r + it;
        }
    }
    const unsigned long long t1 = clock64();
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += a[u];
    if (acc == 0x12345) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP, bool DEP>
void run(const char *name)
{
    unsigned long long *d;
    uint32_t *sink;
    cudaMalloc(&d, 8);
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    for (uint32_t pat = 0; pat < 3; ++pat) {
        printf("%-10s %-5s pattern=%s :", name, DEP ? "dep" : "indep", pat == 0 ? "distinct" : pat == 1 ? "same    " : "random  ");
        for (int wps = 1; wps <= 16; wps *= 2) {            // warps per SM sub-partition
            const int threads = 32 * 4 * (wps > 8 ? 8 : wps);
            const int blocks = 148 * (wps > 8 ? wps / 8 : 1);
            k<OP, DEP><<<blocks, threads>>>(pat, iters, d, sink);
            unsigned long long c = 0;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("  w%-2d %7.1f", wps, (double)c / (iters * 8.0));
        }
        printf("   cyc/op (per warp)\n");
    }
    cudaFree(d);
    cudaFree(sink);
}

int main()
{
    run<MATCH, true>("match.any");
    run<MATCH, false>("match.any");
    run<BALLOT, false>("ballot");
    run<SHFL, false>("shfl");
    run<REDUX, false>("redux.add");
    run<ATOMS_OR, false>("atoms.or");
    run<LDS_RAND, false>("lds");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
