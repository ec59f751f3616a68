// In which order does ATOMS.OR serialise the lanes of a warp that hit the same shared-memory word?
// k_compress's phase A can skip the read-back of the table entry when the lanes are taken in ascending order
// (the returned old value then holds exactly the lower lanes of the same byte value).  Prints, for a few
// value patterns, how many of 1e5 warp-wide atomics saw a higher lane served before a lower one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(const uint8_t *vals, int n_chunks, unsigned long long *viol, unsigned long long *total)
{
    __shared__ uint32_t T[256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) T[i] = 0;
    __syncthreads();
    unsigned long long v_cnt = 0, t_cnt = 0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint32_t v = vals[c * 32 + lane];
        const uint32_t old = atomicOr(&T[v], 1u << lane);
        __syncwarp();
        const uint32_t full = T[v];
        if ((old >> lane) != 0u) ++v_cnt;                          // a higher lane went first
        if (old != (full & ((1u << lane) - 1u))) ++t_cnt;          // old differs from "exactly the lower lanes"
        __syncwarp();
        T[v] = 0;
        __syncwarp();
    }
    atomicAdd(viol, v_cnt);
    atomicAdd(total, t_cnt);
}

int main()
{
    const int n_chunks = 100000;
    uint8_t *h = new uint8_t[n_chunks * 32], *d;
    unsigned long long *dv, hv[2];
    cudaMalloc(&d, n_chunks * 32);
    cudaMalloc(&dv, 16);
    const char *names[] = {"random bytes", "16 values", "4 values", "all same", "pairs at distance 1..8"};
    for (int pat = 0; pat < 5; ++pat) {
        uint32_t s = 12345;
        for (int i = 0; i < n_chunks * 32; ++i) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = s >> 16;
            h[i] = pat == 0 ? r & 255 : pat == 1 ? r & 15 : pat == 2 ? r & 3 : pat == 3 ? 7 : ((i & 31) >= 8 && (r & 1)) ? h[i - 1 - (r >> 1) % 8] : r & 255;
        }
        cudaMemcpy(d, h, n_chunks * 32, cudaMemcpyHostToDevice);
        cudaMemset(dv, 0, 16);
        k<<<148, 32>>>(d, n_chunks, dv, dv + 1);
        cudaMemcpy(hv, dv, 16, cudaMemcpyDeviceToHost);
        printf("%-24s higher-lane-first lanes: %llu   old != lower-lane mask: %llu   (of %d lane-atomics)\n", names[pat], hv[0], hv[1],
               n_chunks * 32);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
