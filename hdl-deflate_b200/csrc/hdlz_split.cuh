// hdlz_split.cuh — shared declarations of the two-phase inflater (hdlz_inflate_split.cu) and its
// launcher (hdlz_inflate_lanes.cu).
#pragma once

#include "hdlz_common.cuh"

namespace hdlz {

constexpr uint32_t kSplitMaxOut = 32768;       // phase 2 stages a whole stream's output in shared memory (OBSIZE)

// per resident decoder thread, global memory: what the rare paths of phase 1 need (code lengths while a block
// header is parsed, sorted symbols + resume point of the bit-serial decode of codes longer than the tables)
struct alignas(8) SplitScratch {          // sorted_l / sorted_d are read as 32-bit words by the decoder
    uint16_t sorted_l[288];
    uint16_t sorted_d[32];
    uint16_t resume_l[2];
    uint16_t resume_d[2];
    uint8_t lens[320];
};

// token words / literal words one stream of at most out_cap output bytes can produce: a copy covers at least
// 3 bytes, a literal-only token at least 252 literals, plus the closing token
__host__ __device__ inline uint32_t split_tokcap(uint32_t out_cap) { return out_cap / 3u + out_cap / 252u + 8u; }
__host__ __device__ inline uint32_t split_litcap_words(uint32_t out_cap) { return out_cap / 4u + 2u; }

size_t split_slot_bytes(uint32_t out_cap);
size_t split_scratch_bytes(const hdlz_ctx *ctx);
int launch_inflate_split(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                         const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                         uint32_t *d_out_len, uint32_t *d_status, uint32_t flags, const uint32_t *d_items,
                         const uint32_t *d_item_count, uint32_t max_items, void *pool, void *scratch,
                         unsigned int *queues, cudaStream_t s);

}  // namespace hdlz
