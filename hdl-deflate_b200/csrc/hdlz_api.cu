// hdlz_api.cu — the extern "C" surface declared in include/hdlz.h: contexts, argument
// checking, host-buffer entry points, device-memory helpers.  No CPU codec lives here:
// every compute entry point ends in a kernel launch or fails.

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hdlz_common.cuh"

namespace hdlz {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    const int code = (e == cudaErrorMemoryAllocation) ? HDLZ_ERR_NOMEM
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? HDLZ_ERR_NODEVICE
                                                                                    : HDLZ_ERR_CUDA;
    return set_error(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

int grow_device(void **p, size_t *cap, size_t need);
static int grow(void **p, size_t *cap, size_t need) { return grow_device(p, cap, need); }

int grow_device(void **p, size_t *cap, size_t need)
{
    if (need <= *cap) return HDLZ_SUCCESS;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(p, need);
        want = need;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
    *cap = want;
    return HDLZ_SUCCESS;
}

static int ensure_pipe(hdlz_ctx *ctx)
{
    for (int i = 0; i < ctx->host_pipe; i++)
        if (!ctx->pipe[i]) {
            cudaError_t e = cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate(pipe)");
        }
    return HDLZ_SUCCESS;
}

// streams per chunk of the *_host pipelines: large enough to fill the persistent grids, small
// enough that copy-in, kernels and copy-out of neighbouring chunks overlap
static inline uint64_t host_chunk(uint64_t n, uint64_t bytes_per_item)
{
    uint64_t c = (48ull << 20) / (bytes_per_item ? bytes_per_item : 1);
    if (c < 8192) c = 8192;
    return c < n ? c : n;
}

// the long-stream compress kernel codes with the fixed or the installed tree, the FAST window and the zlib / raw containers
// (HDLZ_NO_LONG: A/B runs of tools/ and tests/)
static inline bool long_stream_ok(const hdlz_ctx *ctx, uint32_t len)
{
    return len >= HDLZ_LONG_STREAM && ctx->window == HDLZ_CWINDOW && ctx->container != HDLZ_CONTAINER_GZIP && !getenv("HDLZ_NO_LONG");
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_ctx(hdlz_ctx *ctx, DeviceGuard &guard)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    cudaError_t e = guard.enter(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return HDLZ_SUCCESS;
}

// every entry point that touches the device: null check, make the context's device current until return
#define HDLZ_ENTER(ctx)                  \
    DeviceGuard guard__;                 \
    int rc = check_ctx((ctx), guard__);  \
    if (rc) return rc

// lengths of a host batch, checked before anything is launched: a bad length would make a kernel read
// past its slot (a sticky illegal-address error that kills the context)
static int check_lengths(const uint32_t *len, uint64_t n, uint64_t stride, bool compress)
{
    if (!len) return HDLZ_SUCCESS;
    for (uint64_t i = 0; i < n; i++) {
        if (stride && len[i] > stride)
            return set_error(HDLZ_ERR_INVALID, "in_len[%llu] = %u exceeds in_stride %llu", (unsigned long long)i, len[i],
                             (unsigned long long)stride);
        if (compress && len[i] >= (1u << HDLZ_LMAX))
            return set_error(HDLZ_ERR_INVALID, "in_len[%llu] = %u does not fit LMAX", (unsigned long long)i, len[i]);
    }
    return HDLZ_SUCCESS;
}

// error exit of a chunked pipeline: nothing may still be copying into the caller's buffers
static int drain(hdlz_ctx *ctx, int rc)
{
    for (int i = 0; i < kHostPipeMax; i++)
        if (ctx->pipe[i]) cudaStreamSynchronize(ctx->pipe[i]);
    return rc;
}

#define HDLZ_CUDA_DRAIN(ctx, call)                                   \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return drain((ctx), cuda_fail(e__, #call)); \
    } while (0)

}  // namespace hdlz

using namespace hdlz;

extern "C" {

int hdlz_version(void) { return HDLZ_VERSION; }

const char *hdlz_last_error(void) { return g_err; }

const char *hdlz_status_name(uint32_t s)
{
    static const char *names[] = {"OK", "SHORT_INPUT", "BAD_BTYPE", "BAD_CODE", "DIST_TOO_FAR",
                                  "TRUNCATED", "OUT_OVERFLOW", "BAD_STORED", "BAD_HEADER", "BAD_ADLER",
                                  "BAD_CRC", "NO_CODE"};
    return s < sizeof(names) / sizeof(names[0]) ? names[s] : "UNKNOWN";
}

int hdlz_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDeviceCount");
        return 0;
    }
    return n;
}

int hdlz_create(int device, hdlz_ctx **out)
{
    if (!out) return set_error(HDLZ_ERR_INVALID, "null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n == 0) return set_error(HDLZ_ERR_NODEVICE, "no CUDA device");
    if (device < 0 || device >= n) return set_error(HDLZ_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    DeviceGuard guard;
    HDLZ_CUDA(guard.enter(device));
    cudaDeviceProp prop;
    HDLZ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_error(HDLZ_ERR_NODEVICE, "device %d is sm_%d%d; libhdlz is built for sm_100a only", device,
                         prop.major, prop.minor);
    hdlz_ctx *c = new hdlz_ctx();
    memset(c, 0, sizeof *c);
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_match = HDLZ_MAX_MATCH;
    c->window = HDLZ_CWINDOW;
    c->host_pipe = kHostPipe;
    if (const char *e = getenv("HDLZ_HOST_PIPE")) {      // tuning knob of the *_host pipelines (tools/, profiles/)
        const int v = atoi(e);
        if (v >= 2 && v <= kHostPipeMax) c->host_pipe = v;
    }
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return cuda_fail(e, "cudaStreamCreate");
    }
    *out = c;
    return HDLZ_SUCCESS;
}

int hdlz_destroy(hdlz_ctx *c)
{
    if (!c) return HDLZ_SUCCESS;
    DeviceGuard guard;
    guard.enter(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    for (int i = 0; i < kHostPipeMax; i++)
        if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
    if (c->d_in) cudaFree(c->d_in);
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_meta) cudaFree(c->d_meta);
    if (c->d_off) cudaFree(c->d_off);
    for (int i = 0; i < 3; i++) {
        if (c->d_workb[i]) cudaFree(c->d_workb[i]);
        if (c->d_split[i]) cudaFree(c->d_split[i]);
        if (c->d_split_scratch[i]) cudaFree(c->d_split_scratch[i]);
        if (c->slot_event[i]) cudaEventDestroy(c->slot_event[i]);
    }
    if (c->h_dyn_seen) cudaFreeHost(c->h_dyn_seen);
    if (c->d_queue) cudaFree(c->d_queue);
    if (c->d_tree) cudaFree(c->d_tree);
    if (c->d_long) cudaFree(c->d_long);
    if (c->d_pack) cudaFree(c->d_pack);
    for (int i = 0; i < 3; i++)
        if (c->d_lane[i]) cudaFree(c->d_lane[i]);
    if (c->h_small) cudaFreeHost(c->h_small);
    delete c;
    return HDLZ_SUCCESS;
}

uint32_t hdlz_compress_bound(uint32_t len) { return compress_bound(len); }

int hdlz_set_match10(hdlz_ctx *ctx, int match10)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    ctx->max_match = match10 ? HDLZ_MAX_MATCH : HDLZ_MAX_MATCH_SHORT;
    return HDLZ_SUCCESS;
}

int hdlz_get_match10(hdlz_ctx *ctx) { return ctx && ctx->max_match == HDLZ_MAX_MATCH ? 1 : 0; }

int hdlz_set_fast(hdlz_ctx *ctx, int fast)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    ctx->window = fast ? HDLZ_CWINDOW : HDLZ_CWINDOW_SLOW;
    return HDLZ_SUCCESS;
}

int hdlz_get_fast(hdlz_ctx *ctx) { return ctx && ctx->window == HDLZ_CWINDOW_SLOW ? 0 : 1; }

int hdlz_set_container(hdlz_ctx *ctx, int container)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    if (container < HDLZ_CONTAINER_ZLIB || container > HDLZ_CONTAINER_GZIP)
        return set_error(HDLZ_ERR_INVALID, "unknown container %d", container);
    ctx->container = (uint32_t)container;
    return HDLZ_SUCCESS;
}

int hdlz_get_container(hdlz_ctx *ctx) { return ctx ? (int)ctx->container : 0; }

uint32_t hdlz_compress_bound_ex(uint32_t len, int container) { return compress_bound(len, (uint32_t)container); }

int hdlz_compress_batch(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                        uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                        uint32_t *d_status, uint64_t n, void *stream)
{
    HDLZ_ENTER(ctx);
    if (n == 0) return HDLZ_SUCCESS;
    if (!d_in || !d_out || !d_out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (!aligned16(d_in) || !aligned16(d_out) || (in_stride & 15) || (out_stride & 15))
        return set_error(HDLZ_ERR_INVALID, "d_in/d_out must be 16-byte aligned and strides multiples of 16");
    if (!d_in_len && (uniform_len > in_stride || uniform_len >= (1u << HDLZ_LMAX)))
        return set_error(HDLZ_ERR_INVALID, "uniform_len %u does not fit in_stride / LMAX", uniform_len);
    // a few long streams of one length: their tiles are spread over the whole grid instead of one warp per stream
    // (same bytes); with more streams than resident warps the stream-per-warp kernel is as busy and has no look-back
    if (!d_in_len && long_stream_ok(ctx, uniform_len) && n < (uint64_t)ctx->sm_count * 32 && !refresh_tree(ctx) &&
        (ctx->tree_set ? tree_bound(ctx, uniform_len) : compress_bound(uniform_len, ctx->container)) <= out_stride)
        return launch_compress_long(ctx, d_in, in_stride, uniform_len, d_out, out_stride, d_out_len, d_status, n,
                                    (cudaStream_t)stream);
    return launch_compress(ctx, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n,
                           (cudaStream_t)stream);
}

int hdlz_decompress_batch(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                          const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                          uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags, void *stream)
{
    HDLZ_ENTER(ctx);
    if (n == 0) return HDLZ_SUCCESS;
    if (!d_in || !d_in_len || !d_out || !d_out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (out_cap > out_stride) return set_error(HDLZ_ERR_INVALID, "out_cap exceeds out_stride");
    if ((reinterpret_cast<uintptr_t>(d_in) & 3u))
        return set_error(HDLZ_ERR_INVALID, "d_in must be 4-byte aligned");
    return launch_inflate(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len, d_status, n,
                          flags, 0, (cudaStream_t)stream);
}

// The *_host pipelines.  A batch is cut into ~48 MB chunks that go round ctx->host_pipe streams: copy-in of one
// chunk, the kernels of the previous and copy-out of the one before overlap.  What the link is given besides the
// payload decides the rate (tools/pcie_pattern.py: the payload of a 2^20-block round trip alone takes 70-72 ms;
// four small copies per chunk add 11 ms, six streams per call instead of two or three another 8): per-stream
// arrays (lengths, offsets, status words) cross once per call, not once per chunk, a stream is given its next
// chunk only when its previous one is done, and the one small copy a chunk of the packed path needs (its packed
// size and offsets) lands in a pinned staging area of the context.
static int copy_in_lengths(hdlz_ctx *ctx, uint32_t *d_len, const uint32_t *in_len, uint64_t *d_off, const uint64_t *in_off, uint64_t n)
{
    cudaStream_t s = ctx->pipe[0];
    if (in_len) HDLZ_CUDA(cudaMemcpyAsync(d_len, in_len, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (in_off) HDLZ_CUDA(cudaMemcpyAsync(d_off, in_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (in_len || in_off) HDLZ_CUDA(cudaStreamSynchronize(s));      // the other streams' kernels read them
    return HDLZ_SUCCESS;
}

static int copy_out_results(hdlz_ctx *ctx, uint32_t *out_len, const uint32_t *d_olen, uint32_t *status, const uint32_t *d_st, uint64_t n)
{
    for (int i = 0; i < ctx->host_pipe; i++) HDLZ_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->pipe[i]));
    cudaStream_t s = ctx->pipe[0];
    HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(out_len, d_olen, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (status) HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(status, d_st, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA_DRAIN(ctx, cudaStreamSynchronize(s));
    return HDLZ_SUCCESS;
}

int hdlz_compress_host(hdlz_ctx *ctx, const uint8_t *in, uint64_t in_stride, const uint32_t *in_len,
                       uint32_t uniform_len, uint8_t *out, uint64_t out_stride, uint32_t *out_len,
                       uint32_t *status, uint64_t n)
{
    HDLZ_ENTER(ctx);
    if (n == 0) return HDLZ_SUCCESS;
    if (!in || !out || !out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if ((in_stride & 15) || (out_stride & 15)) return set_error(HDLZ_ERR_INVALID, "strides must be multiples of 16");
    if ((rc = check_lengths(in_len, n, in_stride, true))) return rc;
    const size_t in_bytes = (size_t)n * in_stride, out_bytes = (size_t)n * out_stride;
    if ((rc = grow((void **)&ctx->d_in, &ctx->d_in_cap, in_bytes))) return rc;
    if ((rc = grow((void **)&ctx->d_out, &ctx->d_out_cap, out_bytes))) return rc;
    if ((rc = grow((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * n * sizeof(uint32_t)))) return rc;
    uint32_t *d_len = ctx->d_meta, *d_olen = ctx->d_meta + n, *d_st = ctx->d_meta + 2 * n;
    if ((rc = ensure_pipe(ctx))) return rc;
    if ((rc = copy_in_lengths(ctx, d_len, in_len, nullptr, nullptr, n))) return rc;
    const uint64_t chunk = host_chunk(n, in_stride + out_stride);
    int k = 0;
    for (uint64_t first = 0; first < n; first += chunk, ++k) {
        const uint64_t m = n - first < chunk ? n - first : chunk;
        cudaStream_t s = ctx->pipe[k % ctx->host_pipe];
        if (k >= ctx->host_pipe) HDLZ_CUDA_DRAIN(ctx, cudaStreamSynchronize(s));
        HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(ctx->d_in + first * in_stride, in + first * in_stride, m * in_stride,
                                  cudaMemcpyHostToDevice, s));
        rc = hdlz_compress_batch(ctx, ctx->d_in + first * in_stride, in_stride, in_len ? d_len + first : nullptr,
                                 uniform_len, ctx->d_out + first * out_stride, out_stride, d_olen + first, d_st + first,
                                 m, s);
        if (rc) return drain(ctx, rc);
        HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(out + first * out_stride, ctx->d_out + first * out_stride, m * out_stride,
                                  cudaMemcpyDeviceToHost, s));
    }
    return copy_out_results(ctx, out_len, d_olen, status, d_st, n);
}

int hdlz_decompress_host(hdlz_ctx *ctx, const uint8_t *in, const uint64_t *in_off, uint64_t in_stride,
                         const uint32_t *in_len, uint8_t *out, uint64_t out_stride, uint32_t out_cap,
                         uint32_t *out_len, uint32_t *status, uint64_t n, uint32_t flags)
{
    HDLZ_ENTER(ctx);
    if (n == 0) return HDLZ_SUCCESS;
    if (!in || !in_len || !out || !out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (out_cap > out_stride) return set_error(HDLZ_ERR_INVALID, "out_cap exceeds out_stride");
    if (!in_off && (rc = check_lengths(in_len, n, in_stride, false))) return rc;
    size_t in_bytes;
    if (in_off) {
        in_bytes = 0;
        for (uint64_t i = 0; i < n; i++) {
            const size_t e = (size_t)in_off[i] + in_len[i];
            if (e > in_bytes) in_bytes = e;
        }
    } else {
        in_bytes = (size_t)n * in_stride;
    }
    const size_t out_bytes = (size_t)n * out_stride;
    if ((rc = grow((void **)&ctx->d_in, &ctx->d_in_cap, in_bytes + 16))) return rc;
    if ((rc = grow((void **)&ctx->d_out, &ctx->d_out_cap, out_bytes))) return rc;
    if ((rc = grow((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * n * sizeof(uint32_t)))) return rc;
    if (in_off && (rc = grow((void **)&ctx->d_off, &ctx->d_off_cap, n * sizeof(uint64_t)))) return rc;
    uint32_t *d_len = ctx->d_meta, *d_olen = ctx->d_meta + n, *d_st = ctx->d_meta + 2 * n;
    if ((rc = ensure_pipe(ctx))) return rc;
    if ((rc = copy_in_lengths(ctx, d_len, in_len, ctx->d_off, in_off, n))) return rc;
    // Packed input is chunked too when its offsets ascend (what hdlz_pack_batch produces); otherwise it goes
    // in one piece.
    bool ascending = true;
    if (in_off)
        for (uint64_t i = 1; i < n && ascending; i++) ascending = in_off[i] >= in_off[i - 1] + in_len[i - 1];
    const uint64_t chunk = (in_off && !ascending) ? n : host_chunk(n, (in_off ? in_bytes / n + 1 : in_stride) + out_stride);
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    int k = 0;
    for (uint64_t first = 0; first < n; first += chunk, ++k) {
        const uint64_t m = n - first < chunk ? n - first : chunk;
        cudaStream_t s = ctx->pipe[k % ctx->host_pipe];
        // a stream gets its next chunk when its previous one is done: a second call running beside this one (another
        // context, another host thread) finds room in the copy engines' queues between ours instead of behind them
        if (k >= ctx->host_pipe) HDLZ_CUDA_DRAIN(ctx, cudaStreamSynchronize(s));
        if (in_off) {
            const uint64_t lo = (nchunks == 1 ? 0 : in_off[first]) & ~(uint64_t)15;
            const uint64_t hi = nchunks == 1 ? in_bytes : in_off[first + m - 1] + in_len[first + m - 1];
            HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(ctx->d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, s));
        } else {
            HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(ctx->d_in + first * in_stride, in + first * in_stride, m * in_stride,
                                  cudaMemcpyHostToDevice, s));
        }
        rc = launch_inflate(ctx, in_off ? ctx->d_in : ctx->d_in + first * in_stride, in_off ? ctx->d_off + first : nullptr,
                            in_stride, d_len + first, ctx->d_out + first * out_stride, out_stride, out_cap,
                            d_olen + first, d_st + first, m, flags, k % 3, s);
        if (rc) return drain(ctx, rc);
        HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(out + first * out_stride, ctx->d_out + first * out_stride, m * out_stride,
                                  cudaMemcpyDeviceToHost, s));
    }
    return copy_out_results(ctx, out_len, d_olen, status, d_st, n);
}

int hdlz_pack_batch(hdlz_ctx *ctx, const uint8_t *d_slots, uint64_t stride, const uint32_t *d_len, uint8_t *d_packed,
                    uint64_t *d_off, uint64_t *d_total, uint64_t n, void *stream)
{
    HDLZ_ENTER(ctx);
    if (n == 0) return HDLZ_SUCCESS;
    if (!d_slots || !d_len || !d_packed || !d_off || !d_total) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if ((reinterpret_cast<uintptr_t>(d_slots) & 3u) || (reinterpret_cast<uintptr_t>(d_packed) & 3u) || (stride & 3u))
        return set_error(HDLZ_ERR_INVALID, "slots, packed buffer and stride must be 4-byte aligned");
    return launch_pack(ctx, d_slots, stride, d_len, d_packed, d_off, d_total, n, (cudaStream_t)stream);
}

int hdlz_compress_host_packed(hdlz_ctx *ctx, const uint8_t *in, uint64_t in_stride, const uint32_t *in_len,
                              uint32_t uniform_len, uint8_t *out, uint64_t out_cap, uint64_t *out_off,
                              uint32_t *out_len, uint32_t *status, uint64_t n, uint64_t *out_total)
{
    HDLZ_ENTER(ctx);
    if (out_total) *out_total = 0;
    if (n == 0) return HDLZ_SUCCESS;
    if (!in || !out || !out_off || !out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (in_stride & 15) return set_error(HDLZ_ERR_INVALID, "in_stride must be a multiple of 16");
    if ((rc = check_lengths(in_len, n, in_stride, true))) return rc;
    if (!in_len && uniform_len > in_stride) return set_error(HDLZ_ERR_INVALID, "uniform_len exceeds in_stride");
    uint32_t maxlen = uniform_len;
    if (in_len) {
        maxlen = 0;
        for (uint64_t i = 0; i < n; i++) maxlen = in_len[i] > maxlen ? in_len[i] : maxlen;
    }
    if ((rc = refresh_tree(ctx))) return rc;
    const uint64_t slot = ctx->tree_set ? tree_bound(ctx, maxlen) : compress_bound(maxlen, ctx->container);
    const uint64_t chunk = host_chunk(n, in_stride + slot);
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    // per chunk c, contiguous on the device and in the pinned staging area: [packed size | chunk-local offsets]
    // at word first(c) + c
    const size_t meta_words = (size_t)n + nchunks;
    if ((rc = grow((void **)&ctx->d_in, &ctx->d_in_cap, (size_t)n * in_stride))) return rc;
    if ((rc = grow((void **)&ctx->d_out, &ctx->d_out_cap, (size_t)n * slot))) return rc;
    if ((rc = grow((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * n * sizeof(uint32_t)))) return rc;
    if ((rc = grow((void **)&ctx->d_pack, &ctx->d_pack_cap, (size_t)n * slot + 16 + meta_words * sizeof(uint64_t)))) return rc;
    if (meta_words > ctx->h_small_cap) {
        if (ctx->h_small) cudaFreeHost(ctx->h_small);
        ctx->h_small = nullptr;
        ctx->h_small_cap = 0;
        HDLZ_CUDA(cudaMallocHost((void **)&ctx->h_small, (meta_words + meta_words / 4 + 64) * sizeof(uint64_t)));
        ctx->h_small_cap = meta_words + meta_words / 4 + 64;
    }
    if ((rc = ensure_pipe(ctx))) return rc;
    uint32_t *d_len = ctx->d_meta, *d_olen = ctx->d_meta + n, *d_st = ctx->d_meta + 2 * n;
    uint8_t *d_packed = ctx->d_pack;                                               // chunk c packs into its slot range
    uint64_t *d_meta2 = reinterpret_cast<uint64_t *>(ctx->d_pack + (((size_t)n * slot + 15) & ~(size_t)15));
    uint64_t *h_meta2 = ctx->h_small;
    if ((rc = copy_in_lengths(ctx, d_len, in_len, nullptr, nullptr, n))) return rc;
    uint64_t base = 0;
    // stage 1 (copy in, compress, pack, the chunk's size and offsets out) of chunk c is enqueued kLag chunks ahead
    // of stage 2 (packed bytes out), which needs the chunk's packed size on the host
    const uint64_t kLag = (uint64_t)ctx->host_pipe - 1;
    for (uint64_t c = 0; c < nchunks + kLag; ++c) {
        if (c < nchunks) {
            const uint64_t first = c * chunk, m = n - first < chunk ? n - first : chunk;
            cudaStream_t s = ctx->pipe[c % ctx->host_pipe];
            HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(ctx->d_in + first * in_stride, in + first * in_stride, m * in_stride,
                                      cudaMemcpyHostToDevice, s));
            rc = hdlz_compress_batch(ctx, ctx->d_in + first * in_stride, in_stride, in_len ? d_len + first : nullptr,
                                     uniform_len, ctx->d_out + first * slot, slot, d_olen + first, d_st + first, m, s);
            if (rc) return drain(ctx, rc);
            uint64_t *dm = d_meta2 + first + c;
            rc = launch_pack(ctx, ctx->d_out + first * slot, slot, d_olen + first, d_packed + first * slot, dm + 1, dm, m, s);
            if (rc) return drain(ctx, rc);
            HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(h_meta2 + first + c, dm, (m + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        }
        if (c >= kLag) {
            const uint64_t cc = c - kLag, first = cc * chunk, m = n - first < chunk ? n - first : chunk;
            cudaStream_t s = ctx->pipe[cc % ctx->host_pipe];
            HDLZ_CUDA_DRAIN(ctx, cudaStreamSynchronize(s));
            const uint64_t *hm = h_meta2 + first + cc;
            const uint64_t tot = hm[0];
            if (base + tot > out_cap)
                return drain(ctx, set_error(HDLZ_ERR_INVALID, "packed output needs more than out_cap = %llu bytes",
                                            (unsigned long long)out_cap));
            HDLZ_CUDA_DRAIN(ctx, cudaMemcpyAsync(out + base, d_packed + first * slot, tot, cudaMemcpyDeviceToHost, s));
            for (uint64_t i = 0; i < m; i++) out_off[first + i] = hm[1 + i] + base;     // chunk-local -> global offsets
            base += tot;
        }
    }
    if ((rc = copy_out_results(ctx, out_len, d_olen, status, d_st, n))) return rc;
    if (out_total) *out_total = base;
    return HDLZ_SUCCESS;
}

int hdlz_compress_stream(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                         uint32_t *out_len, uint32_t *status)
{
    HDLZ_ENTER(ctx);
    if (!in || !out || !out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (len >= (1u << HDLZ_LMAX)) return set_error(HDLZ_ERR_INVALID, "stream longer than 2^LMAX");
    *out_len = 0;
    const size_t in_slot = ((size_t)len + 15) & ~(size_t)15;
    if ((rc = refresh_tree(ctx))) return rc;
    const size_t out_slot = ctx->tree_set ? tree_bound(ctx, len) : compress_bound(len, ctx->container);
    if ((rc = grow((void **)&ctx->d_in, &ctx->d_in_cap, in_slot + 16))) return rc;
    if ((rc = grow((void **)&ctx->d_out, &ctx->d_out_cap, out_slot))) return rc;
    if ((rc = grow((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * sizeof(uint32_t)))) return rc;
    cudaStream_t s = ctx->stream;
    HDLZ_CUDA(cudaMemcpyAsync(ctx->d_in, in, len, cudaMemcpyHostToDevice, s));
    // (a stream of HDLZ_LONG_STREAM bytes or more goes over the whole grid instead of one warp: hdlz_compress_batch)
    rc = hdlz_compress_batch(ctx, ctx->d_in, in_slot ? in_slot : 16, nullptr, len, ctx->d_out, out_slot, ctx->d_meta + 1,
                             ctx->d_meta + 2, 1, s);
    if (rc) return rc;
    uint32_t meta[2] = {0, 0};
    HDLZ_CUDA(cudaMemcpyAsync(meta, ctx->d_meta + 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA(cudaStreamSynchronize(s));
    uint32_t st = meta[1];
    if (st == HDLZ_OK && meta[0] > out_cap) st = HDLZ_ST_OUT_OVERFLOW;
    if (st == HDLZ_OK) {
        HDLZ_CUDA(cudaMemcpyAsync(out, ctx->d_out, meta[0], cudaMemcpyDeviceToHost, s));
        HDLZ_CUDA(cudaStreamSynchronize(s));
        *out_len = meta[0];
    }
    if (status) *status = st;
    return HDLZ_SUCCESS;
}

int hdlz_decompress_stream(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                           uint32_t *out_len, uint32_t *status, uint32_t flags)
{
    HDLZ_ENTER(ctx);
    if (!in || !out_len || (!out && out_cap)) return set_error(HDLZ_ERR_INVALID, "null buffer");
    *out_len = 0;
    const size_t out_slot = ((size_t)out_cap + 15) & ~(size_t)15;
    if ((rc = grow((void **)&ctx->d_in, &ctx->d_in_cap, (size_t)len + 16))) return rc;
    if ((rc = grow((void **)&ctx->d_out, &ctx->d_out_cap, out_slot + 16))) return rc;
    if ((rc = grow((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * sizeof(uint32_t)))) return rc;
    cudaStream_t s = ctx->stream;
    HDLZ_CUDA(cudaMemcpyAsync(ctx->d_in, in, len, cudaMemcpyHostToDevice, s));
    HDLZ_CUDA(cudaMemcpyAsync(ctx->d_meta, &len, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    rc = hdlz_decompress_batch(ctx, ctx->d_in, nullptr, 0, ctx->d_meta, ctx->d_out, out_slot, out_cap, ctx->d_meta + 1,
                               ctx->d_meta + 2, 1, flags, s);
    if (rc) return rc;
    uint32_t meta[2] = {0, 0};
    HDLZ_CUDA(cudaMemcpyAsync(meta, ctx->d_meta + 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA(cudaStreamSynchronize(s));
    if (meta[1] == HDLZ_OK && meta[0]) {
        HDLZ_CUDA(cudaMemcpyAsync(out, ctx->d_out, meta[0], cudaMemcpyDeviceToHost, s));
        HDLZ_CUDA(cudaStreamSynchronize(s));
    }
    *out_len = meta[1] == HDLZ_OK ? meta[0] : 0;
    if (status) *status = meta[1];
    return HDLZ_SUCCESS;
}

// ---- compress stream fed in pieces (see include/hdlz.h) -------------------------------------------------
struct hdlz_cstream {
    hdlz_ctx *ctx;
    uint8_t *d_buf[2];        // input window on the device, ping-pong: bytes [base, received) of the stream
    int cur;
    uint32_t base;            // stream position of d_buf[cur][0] (a multiple of 16)
    uint32_t received;        // stream bytes fed so far
    uint32_t t0;              // next tile to encode (mirror of the device state)
    uint8_t *d_out;           // output of one launch
    uint32_t *d_small;        // out_len | status
    hdlz::StreamCtl *d_ctl;
    unsigned long long *d_queue;
    hdlz::StreamCtl h_ctl;
    bool finished;
};

namespace {
constexpr uint32_t kCsTile = 1024;                      // tile of k_compress
constexpr uint32_t kCsBuf = (1u << 20) + 4096;          // input window: a piece of up to 1 MiB per launch
constexpr uint32_t kCsOut = (kCsBuf / 8) * 9 + 4096;    // nine bits per input byte at most

int cstream_run(hdlz_cstream *st, bool closing, uint8_t *out, uint32_t out_cap, uint32_t *produced)
{
    hdlz_ctx *ctx = st->ctx;
    cudaStream_t s = ctx->stream;
    // tiles whose result no later byte can change: every `di < isize - k` guard of a position in the tile is
    // decided once 34 bytes past the tile have arrived (deflate.py:913-952, 975-977)
    uint32_t t_end = st->t0;
    if (!closing) {
        if (st->received >= kCsTile + 34u) t_end = (st->received - 34u) / kCsTile * kCsTile;
        if (t_end <= st->t0) return HDLZ_SUCCESS;
    }
    st->h_ctl.t0 = st->t0;
    st->h_ctl.t_end = t_end;
    st->h_ctl.final = closing ? 1u : 0u;
    st->h_ctl.out_words = 0;
    HDLZ_CUDA(cudaMemcpyAsync(st->d_ctl, &st->h_ctl, sizeof(hdlz::StreamCtl), cudaMemcpyHostToDevice, s));
    HDLZ_CUDA(cudaMemsetAsync(st->d_queue, 0, sizeof(unsigned long long), s));
    HDLZ_CUDA(cudaMemsetAsync(st->d_small, 0, 2 * sizeof(uint32_t), s));
    // a piece of many tiles goes over the whole grid (k_compress<.., kLong> with the state in *d_ctl), a short one to one warp
    const uint64_t tiles = closing ? ((uint64_t)(st->received - st->t0) + kCsTile - 1) / kCsTile : (t_end - st->t0) / kCsTile;
    const bool spread = tiles >= HDLZ_LONG_STREAM / kCsTile && !getenv("HDLZ_NO_LONG");
    int rc = spread ? launch_compress_piece(ctx, st->d_buf[st->cur] - st->base, st->received, tiles, st->d_out, kCsOut, st->d_small,
                                            st->d_small + 1, st->d_ctl, s)
                    : launch_compress_stream(ctx, st->d_buf[st->cur] - st->base, st->received, st->d_out, st->d_small, st->d_small + 1,
                                             st->d_ctl, st->d_queue, s);
    if (rc) return rc;
    uint32_t small[2] = {0, 0};
    HDLZ_CUDA(cudaMemcpyAsync(&st->h_ctl, st->d_ctl, sizeof(hdlz::StreamCtl), cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA(cudaMemcpyAsync(small, st->d_small, sizeof small, cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA(cudaStreamSynchronize(s));
    if (spread && !closing) {
        // the partial word after the whole ones stays in the launch's output: it is the next launch's `pw`
        HDLZ_CUDA(cudaMemcpyAsync(&st->h_ctl.pw, st->d_out + 4u * (size_t)st->h_ctl.out_words, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        HDLZ_CUDA(cudaStreamSynchronize(s));
    }
    const uint32_t nbytes = closing ? small[0] : 4u * st->h_ctl.out_words;
    if (nbytes > out_cap) return set_error(HDLZ_ERR_INVALID, "stream output needs %u bytes, out_cap is %u", nbytes, out_cap);
    if (nbytes) HDLZ_CUDA(cudaMemcpyAsync(out, st->d_out, nbytes, cudaMemcpyDeviceToHost, s));
    HDLZ_CUDA(cudaStreamSynchronize(s));
    *produced = nbytes;
    if (!closing) {
        st->t0 = t_end;
        // slide the window: keep the 32 bytes before the next tile (the search window, deflate.py:442-453)
        const uint32_t nb = (st->t0 - 32u) & ~15u;
        if (nb > st->base) {
            HDLZ_CUDA(cudaMemcpyAsync(st->d_buf[st->cur ^ 1], st->d_buf[st->cur] + (nb - st->base), st->received - nb,
                                      cudaMemcpyDeviceToDevice, s));
            st->cur ^= 1;
            st->base = nb;
        }
    }
    return HDLZ_SUCCESS;
}
}  // namespace

int hdlz_cstream_begin(hdlz_ctx *ctx, hdlz_cstream **out)
{
    HDLZ_ENTER(ctx);
    if (!out) return set_error(HDLZ_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (ctx->container == HDLZ_CONTAINER_GZIP)
        return set_error(HDLZ_ERR_INVALID, "streams fed in pieces write the zlib or the raw container (the gzip CRC-32 is taken over the whole input)");
    if (ctx->tree_set) return set_error(HDLZ_ERR_INVALID, "streams fed in pieces code with the fixed tree (hdlz_set_tree is for whole-stream and batch calls)");
    if (ctx->window != HDLZ_CWINDOW) return set_error(HDLZ_ERR_INVALID, "streams fed in pieces use the FAST engine (CWINDOW = 32)");
    hdlz_cstream *st = new hdlz_cstream();
    memset(st, 0, sizeof *st);
    st->ctx = ctx;
    cudaError_t e = cudaMalloc((void **)&st->d_buf[0], kCsBuf + 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_buf[1], kCsBuf + 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_out, kCsOut);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_small, 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_ctl, sizeof(hdlz::StreamCtl));
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_queue, sizeof(unsigned long long));
    if (e != cudaSuccess) {
        hdlz_cstream_end(st);
        return cuda_fail(e, "cudaMalloc(stream)");
    }
    *out = st;
    return HDLZ_SUCCESS;
}

int hdlz_cstream_feed(hdlz_cstream *st, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap, uint32_t *out_len,
                      uint32_t *in_progress)
{
    if (!st) return set_error(HDLZ_ERR_INVALID, "null stream");
    HDLZ_ENTER(st->ctx);
    if ((!in && len) || !out_len || (!out && out_cap)) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (st->finished) return set_error(HDLZ_ERR_INVALID, "stream already finished");
    if ((uint64_t)st->received + len >= (1u << HDLZ_LMAX)) return set_error(HDLZ_ERR_INVALID, "stream longer than 2^LMAX");
    *out_len = 0;
    uint32_t done = 0;
    while (done < len) {
        // as much of the piece as the window holds, then everything that has become final
        const uint32_t room = kCsBuf - (st->received - st->base);
        const uint32_t n = len - done < room ? len - done : room;
        HDLZ_CUDA(cudaMemcpyAsync(st->d_buf[st->cur] + (st->received - st->base), in + done, n, cudaMemcpyHostToDevice,
                                  st->ctx->stream));
        st->received += n;
        done += n;
        uint32_t produced = 0;
        rc = cstream_run(st, false, out + *out_len, out_cap - *out_len, &produced);
        if (rc) return rc;
        *out_len += produced;
    }
    HDLZ_CUDA(cudaStreamSynchronize(st->ctx->stream));       // `in` may be reused by the caller
    if (in_progress) *in_progress = st->t0;
    return HDLZ_SUCCESS;
}

int hdlz_cstream_finish(hdlz_cstream *st, uint8_t *out, uint32_t out_cap, uint32_t *out_len, uint32_t *status)
{
    if (!st) return set_error(HDLZ_ERR_INVALID, "null stream");
    HDLZ_ENTER(st->ctx);
    if (!out_len || (!out && out_cap)) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (st->finished) return set_error(HDLZ_ERR_INVALID, "stream already finished");
    *out_len = 0;
    st->finished = true;
    if (st->received < HDLZ_MIN_INPUT) {                       // the reference never starts (deflate.py:429-432)
        if (status) *status = HDLZ_ST_SHORT_INPUT;
        return HDLZ_SUCCESS;
    }
    uint32_t produced = 0;
    rc = cstream_run(st, true, out, out_cap, &produced);
    if (rc) return rc;
    *out_len = produced;
    if (status) *status = HDLZ_OK;
    return HDLZ_SUCCESS;
}

int hdlz_cstream_end(hdlz_cstream *st)
{
    if (!st) return HDLZ_SUCCESS;
    DeviceGuard guard;
    guard.enter(st->ctx->device);
    cudaStreamSynchronize(st->ctx->stream);
    for (int i = 0; i < 2; i++)
        if (st->d_buf[i]) cudaFree(st->d_buf[i]);
    if (st->d_out) cudaFree(st->d_out);
    if (st->d_small) cudaFree(st->d_small);
    if (st->d_ctl) cudaFree(st->d_ctl);
    if (st->d_queue) cudaFree(st->d_queue);
    delete st;
    return HDLZ_SUCCESS;
}

// ---- decompress stream fed in pieces (see include/hdlz.h) -----------------------------------------------
struct hdlz_dstream {
    hdlz_ctx *ctx;
    uint8_t *d_in;            // every byte fed so far
    size_t in_cap;
    uint32_t received;
    uint8_t *d_out;           // the stream's whole output: also the window of the back-references
    uint32_t out_cap;
    uint32_t flags;
    hdlz::InflateCtl *d_ctl;
    hdlz::InflateCtl h_ctl;   // mirror of the device record after the last launch
    uint32_t delivered;       // output bytes handed to the caller so far
    bool final_run;           // the closing launch has run
};

namespace {
// runs the decoder over what has arrived and hands the caller up to out_cap of the bytes not yet delivered
int dstream_run(hdlz_dstream *st, bool final_input, uint8_t *out, uint32_t out_cap, uint32_t *out_len)
{
    hdlz_ctx *ctx = st->ctx;
    cudaStream_t s = ctx->stream;
    if (!st->h_ctl.done && !st->final_run) {
        int rc = launch_inflate_stream(ctx, st->d_in, st->received, final_input, st->d_out, st->out_cap, st->flags, st->d_ctl, s);
        if (rc) return rc;
        HDLZ_CUDA(cudaMemcpyAsync(&st->h_ctl, st->d_ctl, sizeof(hdlz::InflateCtl), cudaMemcpyDeviceToHost, s));
        HDLZ_CUDA(cudaStreamSynchronize(s));
        if (final_input) st->final_run = true;
    }
    const uint32_t have = st->h_ctl.o - st->delivered;
    const uint32_t n = have < out_cap ? have : out_cap;
    if (n) {
        HDLZ_CUDA(cudaMemcpyAsync(out, st->d_out + st->delivered, n, cudaMemcpyDeviceToHost, s));
        HDLZ_CUDA(cudaStreamSynchronize(s));
        st->delivered += n;
    }
    *out_len = n;
    return HDLZ_SUCCESS;
}
}  // namespace

int hdlz_dstream_begin(hdlz_ctx *ctx, uint32_t max_out, uint32_t flags, hdlz_dstream **out)
{
    HDLZ_ENTER(ctx);
    if (!out) return set_error(HDLZ_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (max_out >= (1u << HDLZ_LMAX)) return set_error(HDLZ_ERR_INVALID, "max_out does not fit LMAX");
    hdlz_dstream *st = new hdlz_dstream();
    memset(st, 0, sizeof *st);
    st->ctx = ctx;
    st->out_cap = max_out;
    st->flags = flags & (HDLZ_F_VERIFY_HEADER | HDLZ_F_VERIFY_ADLER | HDLZ_F_RAW | HDLZ_F_GZIP);
    st->in_cap = 1u << 16;
    cudaError_t e = cudaMalloc((void **)&st->d_in, st->in_cap + 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_out, (size_t)max_out + 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_ctl, sizeof(hdlz::InflateCtl));
    if (e == cudaSuccess) e = cudaMemsetAsync(st->d_ctl, 0, sizeof(hdlz::InflateCtl), ctx->stream);
    if (e != cudaSuccess) {
        hdlz_dstream_end(st);
        return cuda_fail(e, "cudaMalloc(stream)");
    }
    *out = st;
    return HDLZ_SUCCESS;
}

int hdlz_dstream_feed(hdlz_dstream *st, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap, uint32_t *out_len,
                      uint32_t *in_progress)
{
    if (!st) return set_error(HDLZ_ERR_INVALID, "null stream");
    HDLZ_ENTER(st->ctx);
    if ((!in && len) || !out_len || (!out && out_cap)) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (st->final_run) return set_error(HDLZ_ERR_INVALID, "stream already finished");
    if ((uint64_t)st->received + len >= (1u << HDLZ_LMAX)) return set_error(HDLZ_ERR_INVALID, "stream longer than 2^LMAX");
    cudaStream_t s = st->ctx->stream;
    *out_len = 0;
    if ((size_t)st->received + len > st->in_cap) {            // grow: the bytes so far move to the new buffer
        size_t cap = st->in_cap;
        while (cap < (size_t)st->received + len) cap *= 2;
        uint8_t *nb = nullptr;
        HDLZ_CUDA(cudaMalloc((void **)&nb, cap + 16));
        cudaError_t e = cudaMemcpyAsync(nb, st->d_in, st->received, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            cudaFree(nb);
            return cuda_fail(e, "stream input move");
        }
        cudaFree(st->d_in);
        st->d_in = nb;
        st->in_cap = cap;
    }
    if (len) HDLZ_CUDA(cudaMemcpyAsync(st->d_in + st->received, in, len, cudaMemcpyHostToDevice, s));
    st->received += len;
    rc = dstream_run(st, false, out, out_cap, out_len);      // synchronises: `in` may be reused by the caller
    if (rc) return rc;
    if (in_progress) *in_progress = st->h_ctl.started ? (uint32_t)(st->h_ctl.bitpos >> 3) : 0u;
    return HDLZ_SUCCESS;
}

int hdlz_dstream_finish(hdlz_dstream *st, uint8_t *out, uint32_t out_cap, uint32_t *out_len, uint32_t *remaining,
                        uint32_t *status)
{
    if (!st) return set_error(HDLZ_ERR_INVALID, "null stream");
    HDLZ_ENTER(st->ctx);
    if (!out_len || (!out && out_cap)) return set_error(HDLZ_ERR_INVALID, "null buffer");
    *out_len = 0;
    rc = dstream_run(st, true, out, out_cap, out_len);
    if (rc) return rc;
    if (remaining) *remaining = st->h_ctl.o - st->delivered;
    if (status) *status = st->h_ctl.status;
    return HDLZ_SUCCESS;
}

int hdlz_dstream_end(hdlz_dstream *st)
{
    if (!st) return HDLZ_SUCCESS;
    DeviceGuard guard;
    guard.enter(st->ctx->device);
    cudaStreamSynchronize(st->ctx->stream);
    if (st->d_in) cudaFree(st->d_in);
    if (st->d_out) cudaFree(st->d_out);
    if (st->d_ctl) cudaFree(st->d_ctl);
    delete st;
    return HDLZ_SUCCESS;
}

int hdlz_dev_alloc(hdlz_ctx *ctx, size_t bytes, void **d_ptr)
{
    HDLZ_ENTER(ctx);
    if (!d_ptr) return set_error(HDLZ_ERR_INVALID, "null output pointer");
    HDLZ_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return HDLZ_SUCCESS;
}

int hdlz_dev_free(hdlz_ctx *ctx, void *d_ptr)
{
    HDLZ_ENTER(ctx);
    HDLZ_CUDA(cudaFree(d_ptr));
    return HDLZ_SUCCESS;
}

int hdlz_host_alloc_pinned(hdlz_ctx *ctx, size_t bytes, void **h_ptr)
{
    HDLZ_ENTER(ctx);
    if (!h_ptr) return set_error(HDLZ_ERR_INVALID, "null output pointer");
    HDLZ_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16));
    return HDLZ_SUCCESS;
}

int hdlz_host_free_pinned(hdlz_ctx *ctx, void *h_ptr)
{
    HDLZ_ENTER(ctx);
    HDLZ_CUDA(cudaFreeHost(h_ptr));
    return HDLZ_SUCCESS;
}

int hdlz_copy_h2d(hdlz_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    HDLZ_ENTER(ctx);
    HDLZ_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return HDLZ_SUCCESS;
}

int hdlz_copy_d2h(hdlz_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    HDLZ_ENTER(ctx);
    HDLZ_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return HDLZ_SUCCESS;
}

int hdlz_stream_sync(hdlz_ctx *ctx, void *stream)
{
    HDLZ_ENTER(ctx);
    HDLZ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return HDLZ_SUCCESS;
}

int hdlz_generate_blocks(hdlz_ctx *ctx, uint8_t *d_out, uint64_t stride, uint32_t len, uint64_t n, uint64_t seed,
                         uint64_t first_block, void *stream)
{
    HDLZ_ENTER(ctx);
    if (!d_out) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (len > stride) return set_error(HDLZ_ERR_INVALID, "len exceeds stride");
    return launch_generate(ctx, d_out, stride, len, n, seed, first_block, (cudaStream_t)stream);
}

uint64_t hdlz_launch_count(hdlz_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
