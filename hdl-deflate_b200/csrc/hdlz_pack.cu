// hdlz_pack.cu — packing of the fixed-stride stream slots a compress batch produces into one
// contiguous buffer + offset array (the layout hdlz_decompress_batch takes with d_in_off).
// Not part of the codec: it exists so that only real stream bytes cross PCIe / NVLink
// (SURVEY.md 8(f) rank 2 "output compaction + block index").  Starts are 4-byte aligned.

#include "hdlz_common.cuh"
#include "hdlz_frame.cuh"

namespace hdlz {
namespace {

// exclusive prefix sum of round4(len[i]) -> off[i] (64-bit), total -> *total.  One CTA.
__global__ void __launch_bounds__(1024)
k_scan_offsets(const uint32_t *__restrict__ len, uint64_t *__restrict__ off, uint64_t *__restrict__ total, uint64_t n)
{
    __shared__ uint64_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = (uint64_t)tid * per, hi = lo + per < n ? lo + per : n;
    uint64_t sum = 0;
    for (uint64_t i = lo; i < hi; ++i) sum += (len[i] + 3u) & ~3u;
    uint64_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(HDLZ_FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(HDLZ_FULL_MASK, w, d);
            if (lane >= d) w += o;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint64_t base = incl - sum + (warp ? s_warp[warp - 1] : 0);
    for (uint64_t i = lo; i < hi; ++i) {
        off[i] = base;
        base += (len[i] + 3u) & ~3u;
    }
    if (tid == 1023) *total = s_warp[31];
}

// one warp per stream: copy round4(len) bytes from its slot to its packed position
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ slots, uint64_t stride, const uint32_t *__restrict__ len,
       const uint64_t *__restrict__ off, uint8_t *__restrict__ packed, uint64_t n)
{
    const uint64_t nw = (uint64_t)gridDim.x * 8;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += nw) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(slots + i * stride);
        uint32_t *dst = reinterpret_cast<uint32_t *>(packed + off[i]);
        const uint32_t words = (len[i] + 3u) >> 2;
        for (uint32_t k = lane; k < words; k += 32) dst[k] = src[k];
    }
}

// gzip container: CRC-32 of every input stream into the trailer slot k_compress left (the eight
// bytes before the end of the stream: CRC-32 | ISIZE).  One thread per stream; only runs when the
// gzip container is selected.
__global__ void __launch_bounds__(128)
k_gzip_crc(const uint8_t *__restrict__ in, uint64_t in_stride, const uint32_t *__restrict__ in_len, uint32_t uniform_len,
           uint8_t *__restrict__ out, uint64_t out_stride, const uint32_t *__restrict__ out_len, uint64_t n)
{
    __shared__ uint32_t s_nib[16];
    if (threadIdx.x < 16) s_nib[threadIdx.x] = crc32_nibble_entry(threadIdx.x);
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t olen = out_len[i];
    if (olen < 18) return;                                   // the stream was not produced (status says why)
    const uint32_t L = in_len ? in_len[i] : uniform_len;
    const uint32_t crc = crc32_bytes(in + i * in_stride, L, s_nib);
    uint8_t *t = out + i * out_stride + olen - 8;
    for (int b = 0; b < 4; ++b) t[b] = (uint8_t)(crc >> (8 * b));
}

}  // namespace

int launch_gzip_trailers(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                         uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, const uint32_t *d_out_len,
                         uint64_t n, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    k_gzip_crc<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, in_stride, d_in_len, uniform_len, d_out, out_stride,
                                                           d_out_len, n);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

int launch_pack(hdlz_ctx *ctx, const uint8_t *d_slots, uint64_t stride, const uint32_t *d_len, uint8_t *d_packed,
                uint64_t *d_off, uint64_t *d_total, uint64_t n, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    k_scan_offsets<<<1, 1024, 0, s>>>(d_len, d_off, d_total, n);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    uint64_t blocks = (n + 7) / 8;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_pack<<<(unsigned)blocks, 256, 0, s>>>(d_slots, stride, d_len, d_off, d_packed, n);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
