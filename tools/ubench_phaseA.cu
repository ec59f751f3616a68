// Developer tool: candidate formulations of the compress kernel's phase A (byte-equality masks
// R[p], bit k <=> x[p-(32-k)] == x[p]) timed in isolation on workload-like bytes, 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_phaseA tools/ubench_phaseA.cu && ./ubench_phaseA
//   V0  MATCH.ANY + 256-entry table of the previous chunk (what k_compress does today)
//   V1  one 256-entry table, ATOMS.OR for the current chunk
//   V2  two 16-entry nibble tables (low / high nibble -> lane mask) filled with ATOMS.OR, double buffered
//   V3  the same nibble tables computed from 8 ballots, no atomics
// All variants must print the same checksum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kTile = 1024, kChunks = kTile / 32, kWarps = 4;

struct WarpS {
    uint8_t in[32 + kTile + 32];
    uint32_t R[kTile + 64];
    uint32_t T[256];
    uint32_t N[2][32];
};

__device__ void fill(uint8_t *b, int n, uint64_t seed)
{
    uint64_t s = seed;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); };
    int pos = 0;
    while (pos < n) {
        if (!pos || (rnd() & 1)) { int k = 1 + rnd() % 8; while (k-- && pos < n) b[pos++] = (uint8_t)rnd(); }
        else { int k = 3 + rnd() % 10, d = 1 + rnd() % (pos < 32 ? pos : 32); while (k-- && pos < n) { b[pos] = b[pos - d]; ++pos; } }
    }
}

template <int V>
__global__ void __launch_bounds__(kWarps * 32, 8) k(int iters, uint32_t *sums)
{
    extern __shared__ uint4 raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpS &ws = reinterpret_cast<WarpS *>(raw)[warp];
    if (lane == 0) fill(ws.in, 32 + kTile + 32, 0x9E3779B97F4A7C15ull * (blockIdx.x * kWarps + warp + 1));
    for (int i = lane; i < 256; i += 32) ws.T[i] = 0;
    ws.N[0][lane] = 0;
    ws.N[1][lane] = 0;
    __syncwarp();
    uint32_t *T = ws.T, *Rw = ws.R;
    const uint8_t *in_s = ws.in;
    uint32_t check = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t vprev = 0;
        if (V == 0) {
            uint32_t v1 = in_s[lane], m1 = __match_any_sync(0xFFFFFFFFu, v1);
            uint32_t v2 = in_s[32 + lane], m2 = __match_any_sync(0xFFFFFFFFu, v2);
            for (int c = -1; c <= kChunks; ++c) {
                const int i = 32 * c + lane;
                const uint32_t v = v1, mcur = m1;
                v1 = v2; m1 = m2;
                if (c + 2 <= kChunks) { v2 = in_s[32 + 32 * (c + 2) + lane]; m2 = __match_any_sync(0xFFFFFFFFu, v2); }
                const uint32_t mprev = T[v];
                __syncwarp();
                T[vprev] = 0;
                __syncwarp();
                T[v] = mcur;
                __syncwarp();
                vprev = v;
                if (c >= 0) Rw[i + (i >> 5)] = __funnelshift_r(mprev, mcur, lane);
            }
            T[vprev] = 0;
        } else if (V == 1) {
            for (int c = -1; c <= kChunks; ++c) {
                const int i = 32 * c + lane;
                const uint32_t v = in_s[32 + 32 * c + lane];
                const uint32_t mprev = T[v];
                __syncwarp();
                T[vprev] = 0;
                __syncwarp();
                atomicOr(&T[v], 1u << lane);
                __syncwarp();
                const uint32_t mcur = T[v];
                vprev = v;
                if (c >= 0) Rw[i + (i >> 5)] = __funnelshift_r(mprev, mcur, lane);
            }
            __syncwarp();
            T[vprev] = 0;
        } else if (V == 2) {
            for (int c = -1; c <= kChunks; ++c) {
                const int i = 32 * c + lane;
                const uint32_t v = in_s[32 + 32 * c + lane];
                uint32_t *cur = ws.N[c & 1], *prv = ws.N[(c & 1) ^ 1];
                const uint32_t lo = v & 15u, hi = 16u + (v >> 4);
                cur[lane] = 0;                                  // the table of chunk c-2 is dead
                const uint32_t mprev = prv[lo] & prv[hi];
                __syncwarp();
                atomicOr(&cur[lo], 1u << lane);
                atomicOr(&cur[hi], 1u << lane);
                __syncwarp();
                const uint32_t mcur = cur[lo] & cur[hi];
                if (c >= 0) Rw[i + (i >> 5)] = __funnelshift_r(mprev, mcur, lane);
            }
            __syncwarp();
            ws.N[0][lane] = 0;
            ws.N[1][lane] = 0;
        } else {
            // lane l < 16 owns the low-nibble entry l, lane l >= 16 the high-nibble entry l - 16
            const uint32_t n = lane & 15u;
            const uint32_t c0 = (n & 1u) ? 0u : ~0u, c1 = (n & 2u) ? 0u : ~0u, c2 = (n & 4u) ? 0u : ~0u, c3 = (n & 8u) ? 0u : ~0u;
            const bool upper = lane >= 16;
            for (int c = -1; c <= kChunks; ++c) {
                const int i = 32 * c + lane;
                const uint32_t v = in_s[32 + 32 * c + lane];
                uint32_t *cur = ws.N[c & 1], *prv = ws.N[(c & 1) ^ 1];
                const uint32_t lo = v & 15u, hi = 16u + (v >> 4);
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, v & 1u), b1 = __ballot_sync(0xFFFFFFFFu, v & 2u);
                const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, v & 4u), b3 = __ballot_sync(0xFFFFFFFFu, v & 8u);
                const uint32_t b4 = __ballot_sync(0xFFFFFFFFu, v & 16u), b5 = __ballot_sync(0xFFFFFFFFu, v & 32u);
                const uint32_t b6 = __ballot_sync(0xFFFFFFFFu, v & 64u), b7 = __ballot_sync(0xFFFFFFFFu, v & 128u);
                const uint32_t e = ((upper ? b4 : b0) ^ c0) & ((upper ? b5 : b1) ^ c1) & ((upper ? b6 : b2) ^ c2) &
                                   ((upper ? b7 : b3) ^ c3);
                const uint32_t mprev = prv[lo] & prv[hi];
                __syncwarp();
                cur[lane] = e;
                __syncwarp();
                const uint32_t mcur = cur[lo] & cur[hi];
                if (c >= 0) Rw[i + (i >> 5)] = __funnelshift_r(mprev, mcur, lane);
            }
            __syncwarp();
            ws.N[0][lane] = 0;
            ws.N[1][lane] = 0;
        }
        __syncwarp();
        if (it == 0)
            for (int i = lane; i < kTile + 32; i += 32) check = check * 31u + Rw[i + (i >> 5)];
        __syncwarp();
    }
    check += __shfl_xor_sync(0xFFFFFFFFu, check, 1) * 7u;
    check += __shfl_xor_sync(0xFFFFFFFFu, check, 2) * 13u;
    check += __shfl_xor_sync(0xFFFFFFFFu, check, 4) * 17u;
    check += __shfl_xor_sync(0xFFFFFFFFu, check, 8) * 19u;
    check += __shfl_xor_sync(0xFFFFFFFFu, check, 16) * 23u;
    if (lane == 0) atomicAdd(&sums[V], check);
}

template <int V>
void run(const char *name, uint32_t *d_sums)
{
    const size_t smem = sizeof(WarpS) * kWarps;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 200;
    k<V><<<148 * 8, kWarps * 32, smem>>>(2, d_sums);
    cudaEventRecord(e0);
    k<V><<<148 * 8, kWarps * 32, smem>>>(iters, d_sums);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    uint32_t h[4];
    cudaMemcpy(h, d_sums, 16, cudaMemcpyDeviceToHost);
    const double chunks = 32.0 * iters * (kChunks + 2);      // per SM
    printf("%-28s %8.3f ms  %7.2f cycles per chunk per SM   checksum %08x  (%s)\n", name, ms,
           ms * 1e-3 * 1.965e9 / chunks, h[V], cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    uint32_t *d_sums;
    cudaMalloc(&d_sums, 16);
    cudaMemset(d_sums, 0, 16);
    run<0>("V0 match.any + table", d_sums);
    run<1>("V1 byte table, atoms.or", d_sums);
    run<2>("V2 nibble tables, atoms.or", d_sums);
    run<3>("V3 nibble tables, 8 ballots", d_sums);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
