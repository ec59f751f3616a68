// hdlz_common.cuh — shared declarations of libhdlz (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "hdlz.h"

#define HDLZ_FULL_MASK 0xFFFFFFFFu
// internal routing flags of hdlz_decompress_batch (tests / profiling)
#define HDLZ_F_FORCE_GENERAL 0x100u   /* warp-per-stream kernel only */
#define HDLZ_F_FORCE_LANES 0x200u     /* lane-per-stream kernel first, whatever the batch size */
#define HDLZ_F_PERSISTENT_LANES 0x800u /* fixed/stored lane kernel as a persistent grid (experiment) */
#define HDLZ_F_NO_LANE_SCRATCH 0x400u /* lanes hand dynamic-block streams to the warp-per-stream kernel */
#define HDLZ_F_NO_SPLIT 0x1000u       /* dynamic-block streams stay on the lane-per-stream kernel (no two-phase route) */

namespace hdlz {
// The code a compressor context uses instead of the fixed one (hdlz_tree.cu): token tables of the kernels and the
// bit string every stream starts with (container header, BFINAL / BTYPE = 10, the code description).
constexpr int kTreeLenBase = 288;       // lut[288 + n]: length symbol of a match of 3 + n bytes
constexpr int kTreeLutWords = 296;
constexpr int kTreePrefixWords = 80;    // 80 header bits (gzip) + 17 + 19 * 3 + 316 * 7 bits at most
constexpr int kTreeHistWords = 296;     // hdlz_train_tree: counts in the layout of lut
struct TreeDev {
    uint32_t lut[kTreeLutWords];        // [0..255] literal: code | bits << 16;  [256 + f] distance 32 - f: code and extra
                                        // bits | bits << 24;  [288 + n]: code | bits << 16
    uint32_t dist[256];                 // distance d at [d - 1], as lut[256 + f] (the CWINDOW = 256 kernel)
    uint32_t eob;                       // code | bits << 16
    uint32_t prefix_bits;
    uint32_t worst_bits;                // most bits one input byte can cost
    uint32_t pad;
    uint32_t prefix[kTreePrefixWords];
};
}  // namespace hdlz

struct hdlz_ctx {
    int device;
    int sm_count;
    uint32_t container;  // hdlz_container written by the compressor
    uint32_t max_match;  // longest match of the compressor: 10 (MATCH10, default) or 5 (deflate.py:34-35)
    uint32_t window;     // search window of the compressor: 32 (FAST, default) or 256 (FAST = False, deflate.py:56-59)
    bool wide_attr_set;
    bool stream_attr_set;
    bool tree_attr_set;
    bool long_attr_set, long_tree_attr_set;
    void *d_long;          // look-back arrays of the long-stream compress kernel (one entry per tile)
    size_t d_long_cap;
    // application / trained code (hdlz_set_tree, hdlz_train_tree); tree_set false = the reference's fixed code
    bool tree_set;
    uint32_t tree_container;   // container the prefix of `tree` was built for
    uint8_t tree_lit[286], tree_dist[30];
    hdlz::TreeDev tree;        // host image of *d_tree
    hdlz::TreeDev *d_tree;
    // scratch for the host-buffer entry points (grown on demand, reused)
    uint8_t *d_in;
    size_t d_in_cap;
    uint8_t *d_out;
    size_t d_out_cap;
    uint32_t *d_meta;  // in_len | out_len | status
    size_t d_meta_cap;
    uint64_t *d_off;
    size_t d_off_cap;
    uint8_t *d_pack;   // packed streams + per-chunk offsets/totals of hdlz_compress_host_packed
    size_t d_pack_cap;
    void *d_lane[3];    // per-thread scratch of the lane inflater for dynamic blocks, one per concurrent launch
    size_t d_lane_cap[3];
    uint64_t *h_small;  // pinned: per-chunk packed sizes
    size_t h_small_cap;
    unsigned long long *d_queue;  // work-queue heads of the persistent compress grid, one slot per launch (ring)
    unsigned queue_seq;
    // inflate launches use one of three slots (the *_host pipelines keep three chunks in flight; the batch
    // entry point uses slot 0).  Everything a launch shares with its kernels lives in its slot, and a slot's
    // next launch waits for the event its previous launch recorded: two batches in flight never share
    // hand-over lists, lane scratch or token pools.
    uint32_t *d_workb[3];  // [0] count for the warp kernel, [1] dynamic-block count, [2..3] queue heads of the two-phase
                           // route, [16..] stream ids handed over by the lane kernel (two lists of n)
    size_t d_workb_cap[3];
    cudaEvent_t slot_event[3];
    void *d_split[3];      // token pool of the two-phase dynamic-block route (records | tokens | literals)
    size_t d_split_cap[3];
    void *d_split_scratch[3];
    uint32_t *h_dyn_seen;  // pinned [3]: dynamic-block streams the slot's last launch saw -> sizes the pool of the next one
    bool split_attr_set;
    cudaStream_t stream;  // owned, used by the host-buffer entry points
    cudaStream_t pipe[6];  // owned, created on first use: chunked H2D / kernel / D2H pipeline of the *_host calls
    int host_pipe;         // how many of them the pipelines use (kHostPipe; HDLZ_HOST_PIPE overrides)
    unsigned long long launches;
    // per-context launch state (was function-static: two contexts / threads raced on it)
    bool compress_attr_set;   // cudaFuncSetAttribute of k_compress done for this context's device
    int l2_persist, l2_window, l2_carved;   // persisting-L2 limits of the device, and whether this context carved it
};

namespace hdlz {
constexpr int kHostPipeMax = 6;
constexpr int kHostPipe = 3;   // streams (= chunks in flight) of the *_host pipelines; more only crowd the copy engines' queues (tools/pcie_pattern.py)
// Makes the context's device current for one entry point and restores the caller's device on return
// (an engine on GPU k must not change the calling thread's current device).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int device)
    {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev != device) {
            e = cudaSetDevice(device);
            switched = e == cudaSuccess;
        }
        return e;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};
}  // namespace hdlz

namespace hdlz {

// what a compress stream carries from one launch to the next (device memory; see k_compress<.., true>)
struct StreamCtl {
    uint32_t t0;        // next position to process (a multiple of the tile size)
    uint32_t t_end;     // this launch stops here (not used by the closing launch)
    uint32_t final;     // closing launch: `received` is the stream's true length
    uint32_t carry, adler_a, adler_b, pw, lbit;
    uint32_t out_words; // 32-bit words this launch wrote (the closing launch reports bytes through out_len)
};
// what a decompress stream carries from one launch to the next (device memory; see k_inflate<true>)
struct InflateCtl {
    uint32_t started;      // a resume point is recorded
    uint32_t at_header;    // resume at a block header (bitpos) / inside a Huffman block (tables from hdr_bitpos, then bitpos)
    uint32_t final_blk;    // BFINAL of the block being decoded
    uint32_t o;            // output bytes produced so far
    unsigned long long bitpos, hdr_bitpos;
    uint32_t done;         // the stream has ended (status is final)
    uint32_t status;
};
int launch_inflate_stream(hdlz_ctx *ctx, const uint8_t *d_in, uint32_t received, bool final_input, uint8_t *d_out,
                          uint32_t out_cap, uint32_t flags, InflateCtl *d_ctl, cudaStream_t s);
int launch_compress_stream(hdlz_ctx *ctx, const uint8_t *d_in_virtual, uint32_t received, uint8_t *d_out, uint32_t *d_out_len,
                           uint32_t *d_status, StreamCtl *d_ctl, unsigned long long *d_queue, cudaStream_t s);

// error plumbing (hdlz_api.cu)
int set_error(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define HDLZ_CUDA(call)                                   \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// kernel launchers
int launch_compress(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                    uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                    uint32_t *d_status, uint64_t n, cudaStream_t s);
int launch_compress_long(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, uint32_t len, uint8_t *d_out, uint64_t out_stride,
                         uint32_t *d_out_len, uint32_t *d_status, uint64_t n, cudaStream_t s);
int launch_compress_piece(hdlz_ctx *ctx, const uint8_t *d_in_virtual, uint32_t received, uint64_t n_tiles, uint8_t *d_out,
                          uint64_t out_bytes, uint32_t *d_out_len, uint32_t *d_status, StreamCtl *d_ctl, cudaStream_t s);
int launch_compress_hist(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len, uint32_t uniform_len,
                         uint64_t n, unsigned long long *d_hist, cudaStream_t s);
int refresh_tree(hdlz_ctx *ctx);
uint32_t tree_bound(const hdlz_ctx *ctx, uint32_t len);
int launch_compress_wide(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                         uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                         uint32_t *d_status, uint64_t n, unsigned long long *queue, cudaStream_t s);
int launch_inflate(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                   const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                   uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags, int slot, cudaStream_t s);
int grow_device(void **p, size_t *cap, size_t need);
int launch_pack(hdlz_ctx *ctx, const uint8_t *d_slots, uint64_t stride, const uint32_t *d_len, uint8_t *d_packed,
                uint64_t *d_off, uint64_t *d_total, uint64_t n, cudaStream_t s);
int launch_gzip_trailers(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                         uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, const uint32_t *d_out_len,
                         uint64_t n, cudaStream_t s);
int launch_generate(hdlz_ctx *ctx, uint8_t *d_out, uint64_t stride, uint32_t len, uint64_t n, uint64_t seed,
                    uint64_t first_block, cudaStream_t s);

__host__ __device__ inline uint32_t compress_bound(uint32_t len, uint32_t container = HDLZ_CONTAINER_ZLIB)
{
    // header + ceil((3 + 9*len + 7) / 8) + trailer, rounded up to 16.  zlib: 2 + 4, raw: 0 + 0, gzip: 10 + 8
    const uint64_t frame = container == HDLZ_CONTAINER_GZIP ? 18ull : container == HDLZ_CONTAINER_RAW ? 0ull : 6ull;
    uint64_t b = frame + (3ull + 9ull * len + 7ull + 7ull) / 8ull;
    return (uint32_t)((b + 15ull) & ~15ull);
}

}  // namespace hdlz
