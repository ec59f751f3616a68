// hdlz_workload.cu — device-side generator of the synthetic "random+repeat" blocks of
// BASELINE config 2.  Measurement support, not part of the codec.
// Block b is a pure function of (seed, b); hdl-deflate_b200/workload.py holds the
// bit-identical CPU definition used by the tests.
//
//   repeat until `len` bytes:  r = next()
//     (r & 1) == 0 or pos == 0 : literal run of 1 + ((r >> 1) & 7) bytes, the bytes of next()
//     else                     : copy 3 + ((r >> 1) % 10) bytes from distance
//                                1 + ((r >> 8) % min(32, pos))   (overlap allowed)

#include "hdlz_common.cuh"

namespace hdlz {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(128)
k_generate(uint8_t *out, uint64_t stride, uint32_t len, uint64_t n, uint64_t seed, uint64_t first_block)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    uint8_t *dst = out + b * stride;
    uint64_t s = seed ^ ((first_block + b) * 0xD1342543DE82EF95ull);
    uint32_t pos = 0;
    while (pos < len) {
        s += 0x9E3779B97F4A7C15ull;
        const uint64_t r = mix64(s);
        if ((r & 1) == 0 || pos == 0) {
            uint32_t run = 1 + (uint32_t)((r >> 1) & 7);
            s += 0x9E3779B97F4A7C15ull;
            uint64_t bytes = mix64(s);
            for (; run && pos < len; --run, ++pos, bytes >>= 8) dst[pos] = (uint8_t)bytes;
        } else {
            uint32_t m = 3 + (uint32_t)((r >> 1) % 10);
            const uint32_t lim = pos < 32 ? pos : 32;
            const uint32_t d = 1 + (uint32_t)((r >> 8) % lim);
            for (; m && pos < len; --m, ++pos) dst[pos] = dst[pos - d];
        }
    }
}

}  // namespace

int launch_generate(hdlz_ctx *ctx, uint8_t *d_out, uint64_t stride, uint32_t len, uint64_t n, uint64_t seed,
                    uint64_t first_block, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    const uint64_t blocks = (n + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return set_error(HDLZ_ERR_INVALID, "too many blocks for one launch");
    k_generate<<<(unsigned)blocks, 128, 0, s>>>(d_out, stride, len, n, seed, first_block);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
