// hdlz_tree.cu — the tree a compressor context codes with: the fixed code of RFC 1951 (the reference's only
// mode, deflate.py:112-149 / STATIC :1064-1076) or a code handed in by the application — the "dedicated
// pre-computed Huffman tree" the reference's README names as the next step for data with few byte values
// (README.md:43-45) — or one trained on a batch (hdlz_train_tree: the symbol histogram of the reference's own
// parse, counted on the GPU by the compress kernel, turned into length-limited code lengths here).
//
// Host code of this file: validation of the lengths, canonical codes, the dynamic-block header (HLIT / HDIST /
// HCLEN, run-length coded lengths, code-length code; RFC 1951 3.2.7) as a bit string every stream starts with,
// and the token tables the kernel reads.  The LZ77 parse is untouched: a stream coded with a tree holds the
// same tokens as deflate.py's output for the same input, in a BTYPE = 10 block.
//
// Deterministic by construction (tests/ compares the bytes with an independent restatement in oracle/):
//   lengths from counts   boundary package-merge, items ordered by (weight, leaf before package, symbol)
//   run-length coding     zeros: 18 while >= 11 remain (138 at most), 17 for 3..10, else single zeros;
//                         a non-zero length once, then 16 for 3..6 repeats, else singles; literal/length
//                         and distance lengths coded as ONE sequence
//   code-length code      the same package-merge limited to 7 bits

#include <string.h>

#include <algorithm>
#include <vector>

#include "hdlz_common.cuh"

namespace hdlz {

namespace {

const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                12289, 16385, 24577};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

uint32_t rev(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// Optimal code lengths under a length limit (package-merge in its list form).  `lens` gets 0 for unused symbols.
void limited_lengths(const uint64_t *freq, int n, int maxbits, uint8_t *lens)
{
    std::vector<int> used;
    for (int s = 0; s < n; s++) {
        lens[s] = 0;
        if (freq[s]) used.push_back(s);
    }
    const int m = (int)used.size();
    if (m == 0) return;
    if (m == 1) {                       // a single code: one bit (zlib accepts this incomplete code)
        lens[used[0]] = 1;
        return;
    }
    std::stable_sort(used.begin(), used.end(), [&](int a, int b) { return freq[a] < freq[b]; });   // ties: symbol order
    // level lists: weight and whether the item is a leaf; leaves appear in `used` order inside every list
    std::vector<std::vector<uint64_t>> w(maxbits);
    std::vector<std::vector<uint8_t>> leaf(maxbits);
    for (int i = 0; i < m; i++) {
        w[0].push_back(freq[used[i]]);
        leaf[0].push_back(1);
    }
    for (int l = 1; l < maxbits; l++) {
        const std::vector<uint64_t> &pw = w[l - 1];
        size_t li = 0, pi = 0;
        const size_t npk = pw.size() / 2;
        while (li < (size_t)m || pi < npk) {
            const uint64_t pkw = pi < npk ? pw[2 * pi] + pw[2 * pi + 1] : 0;
            if (li < (size_t)m && (pi >= npk || freq[used[li]] <= pkw)) {     // tie: the leaf goes first
                w[l].push_back(freq[used[li]]);
                leaf[l].push_back(1);
                li++;
            } else {
                w[l].push_back(pkw);
                leaf[l].push_back(0);
                pi++;
            }
        }
    }
    size_t take = 2 * (size_t)m - 2;
    for (int l = maxbits - 1; l >= 0 && take; l--) {
        if (take > w[l].size()) take = w[l].size();
        size_t leaves = 0;
        for (size_t i = 0; i < take; i++) leaves += leaf[l][i];
        for (size_t i = 0; i < leaves; i++) lens[used[i]]++;
        take = 2 * (take - leaves);
    }
}

// canonical codes (RFC 1951 3.2.2), bit-reversed for LSB-first emission like the reference's out_codes
void canonical(const uint8_t *lens, int n, uint32_t *codes)
{
    uint32_t count[16] = {0}, next[16] = {0};
    for (int s = 0; s < n; s++) count[lens[s]]++;
    count[0] = 0;
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + count[l - 1]) << 1;
        next[l] = code;
    }
    for (int s = 0; s < n; s++) codes[s] = lens[s] ? rev(next[lens[s]]++, lens[s]) : 0;
}

// 0 = usable: not over-subscribed; incomplete only as a single one-bit code (zlib's inflate_table rule)
int check_code(const uint8_t *lens, int n, const char *what)
{
    int left = 1, maxlen = 0, used = 0;
    for (int l = 1; l <= 15; l++) {
        int c = 0;
        for (int s = 0; s < n; s++) c += lens[s] == l;
        left = 2 * left - c;
        if (c) maxlen = l;
        used += c;
        if (left < 0) return set_error(HDLZ_ERR_INVALID, "%s code lengths are over-subscribed", what);
    }
    for (int s = 0; s < n; s++)
        if (lens[s] > 15) return set_error(HDLZ_ERR_INVALID, "%s code length %d exceeds 15", what, lens[s]);
    if (used && left > 0 && maxlen != 1) return set_error(HDLZ_ERR_INVALID, "%s code lengths are incomplete", what);
    return HDLZ_SUCCESS;
}

struct BitString {
    std::vector<uint32_t> words;
    uint32_t bits = 0;
    void put(uint32_t v, uint32_t n)
    {
        for (uint32_t i = 0; i < n; i++, bits++) {
            if ((bits >> 5) >= words.size()) words.push_back(0);
            words[bits >> 5] |= ((v >> i) & 1u) << (bits & 31);
        }
    }
};

}  // namespace

static void build_prefix(uint32_t container, const uint8_t *lit, const uint8_t *dist, BitString &bs);

// Fills ctx->tree (host image) from the lengths and uploads it.  The lengths have been validated.
static int install_tree(hdlz_ctx *ctx, const uint8_t *lit, const uint8_t *dist)
{
    TreeDev &t = ctx->tree;
    memset(&t, 0, sizeof t);
    uint32_t lcode[286], dcode[30];
    canonical(lit, 286, lcode);
    canonical(dist, 30, dcode);
    uint32_t maxlit = 0, maxlen = 0, maxdist = 0;
    for (int s = 0; s < 256; s++) {
        t.lut[s] = lcode[s] | ((uint32_t)lit[s] << 16);
        maxlit = std::max<uint32_t>(maxlit, lit[s]);
    }
    for (int n = 0; n < 8; n++) {                         // match length 3 + n: symbol 257 + n, no extra bits (DISTANCE, deflate.py:845-850)
        t.lut[kTreeLenBase + n] = lcode[257 + n] | ((uint32_t)lit[257 + n] << 16);
        maxlen = std::max<uint32_t>(maxlen, lit[257 + n]);
    }
    // distances 1..256 (CWINDOW = 32 uses the first 32): distance code + extra bits as one field (deflate.py:858-874)
    for (int d = 1; d <= 256; d++) {
        int c = 0;
        while (c + 1 < 30 && kDistBase[c + 1] <= d) c++;
        const uint32_t eb = c < 2 ? 0 : (uint32_t)(c >> 1) - 1;
        uint32_t e = 0;
        if (dist[c]) {
            e = (dcode[c] | ((uint32_t)(d - kDistBase[c]) << dist[c])) | ((dist[c] + eb) << 24);
            maxdist = std::max<uint32_t>(maxdist, dist[c] + eb);
        }
        t.dist[d - 1] = e;
        if (d <= 32) t.lut[256 + (32 - d)] = e;          // by mask bit f of the FAST kernel: d = 32 - f
    }
    t.eob = lcode[256] | ((uint32_t)lit[256] << 16);
    // worst case per input byte: a literal, or a third of the dearest three-byte match
    t.worst_bits = std::max(maxlit, (maxlen + maxdist + 2) / 3);
    if (t.worst_bits == 0) t.worst_bits = 1;

    BitString bs;
    build_prefix(ctx->container, lit, dist, bs);
    if (bs.bits >= 32u * kTreePrefixWords) return set_error(HDLZ_ERR_INVALID, "tree description too long (%u bits)", bs.bits);
    t.prefix_bits = bs.bits;
    for (size_t i = 0; i < bs.words.size(); i++) t.prefix[i] = bs.words[i];

    // the kernels of earlier launches may still read the previous tables
    HDLZ_CUDA(cudaDeviceSynchronize());
    if (!ctx->d_tree) HDLZ_CUDA(cudaMalloc((void **)&ctx->d_tree, sizeof(TreeDev)));
    HDLZ_CUDA(cudaMemcpy(ctx->d_tree, &t, sizeof t, cudaMemcpyHostToDevice));
    memcpy(ctx->tree_lit, lit, 286);
    memcpy(ctx->tree_dist, dist, 30);
    ctx->tree_container = ctx->container;
    ctx->tree_set = true;
    return HDLZ_SUCCESS;
}

// What every stream starts with: container header, BFINAL = 1 / BTYPE = 10, the code description (RFC 1951 3.2.7)
static void build_prefix(uint32_t container, const uint8_t *lit, const uint8_t *dist, BitString &bs)
{
    if (container == HDLZ_CONTAINER_ZLIB) {
        bs.put(0x78, 8);                                   // deflate.py:753-757
        bs.put(0x9C, 8);
    } else if (container == HDLZ_CONTAINER_GZIP) {
        const uint8_t hdr[10] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF};
        for (int i = 0; i < 10; i++) bs.put(hdr[i], 8);
    }
    bs.put(1, 1);
    bs.put(2, 2);
    int nlit = 286, ndist = 30;
    while (nlit > 257 && lit[nlit - 1] == 0) nlit--;
    while (ndist > 1 && dist[ndist - 1] == 0) ndist--;
    std::vector<uint8_t> seq(lit, lit + nlit);
    seq.insert(seq.end(), dist, dist + ndist);
    // run-length coding into (symbol, extra value) pairs
    std::vector<std::pair<uint8_t, uint8_t>> rl;
    for (size_t i = 0; i < seq.size();) {
        const uint8_t v = seq[i];
        size_t run = 1;
        while (i + run < seq.size() && seq[i + run] == v) run++;
        i += run;
        if (v == 0) {
            while (run >= 11) {
                const size_t r = std::min<size_t>(run, 138);
                rl.push_back({18, (uint8_t)(r - 11)});
                run -= r;
            }
            if (run >= 3) {
                rl.push_back({17, (uint8_t)(run - 3)});
                run = 0;
            }
            while (run--) rl.push_back({0, 0});
        } else {
            rl.push_back({v, 0});
            run--;
            while (run >= 3) {
                const size_t r = std::min<size_t>(run, 6);
                rl.push_back({16, (uint8_t)(r - 3)});
                run -= r;
            }
            while (run--) rl.push_back({v, 0});
        }
    }
    uint64_t clfreq[19] = {0};
    for (auto &p : rl) clfreq[p.first]++;
    uint8_t cllen[19];
    uint32_t clcode[19];
    limited_lengths(clfreq, 19, 7, cllen);
    canonical(cllen, 19, clcode);
    int ncl = 19;
    while (ncl > 4 && cllen[kClOrder[ncl - 1]] == 0) ncl--;
    bs.put((uint32_t)(nlit - 257), 5);
    bs.put((uint32_t)(ndist - 1), 5);
    bs.put((uint32_t)(ncl - 4), 4);
    for (int i = 0; i < ncl; i++) bs.put(cllen[kClOrder[i]], 3);
    for (auto &p : rl) {
        bs.put(clcode[p.first], cllen[p.first]);
        if (p.first == 16) bs.put(p.second, 2);
        else if (p.first == 17) bs.put(p.second, 3);
        else if (p.first == 18) bs.put(p.second, 7);
    }
}

// the container is part of the prefix: rebuilt when hdlz_set_container changed it after hdlz_set_tree
int refresh_tree(hdlz_ctx *ctx)
{
    if (!ctx->tree_set || ctx->tree_container == ctx->container) return HDLZ_SUCCESS;
    uint8_t lit[286], dist[30];
    memcpy(lit, ctx->tree_lit, 286);
    memcpy(dist, ctx->tree_dist, 30);
    return install_tree(ctx, lit, dist);
}

uint32_t tree_bound(const hdlz_ctx *ctx, uint32_t len)
{
    const TreeDev &t = ctx->tree;
    const uint64_t trailer = ctx->container == HDLZ_CONTAINER_GZIP ? 8 : ctx->container == HDLZ_CONTAINER_RAW ? 0 : 4;
    const uint64_t bits = (uint64_t)t.prefix_bits + (uint64_t)len * t.worst_bits + (t.eob >> 16);
    return (uint32_t)(((bits + 7) / 8 + trailer + 15) & ~15ull);
}

}  // namespace hdlz

using namespace hdlz;

extern "C" {

int hdlz_set_tree(hdlz_ctx *ctx, const uint8_t *lit_len, const uint8_t *dist_len)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    DeviceGuard guard;
    HDLZ_CUDA(guard.enter(ctx->device));
    if (!lit_len && !dist_len) {                 // back to the fixed code
        ctx->tree_set = false;
        return HDLZ_SUCCESS;
    }
    if (!lit_len || !dist_len) return set_error(HDLZ_ERR_INVALID, "both length arrays are needed (or both NULL)");
    int rc;
    if ((rc = check_code(lit_len, 286, "literal/length"))) return rc;
    if ((rc = check_code(dist_len, 30, "distance"))) return rc;
    if (lit_len[256] == 0) return set_error(HDLZ_ERR_INVALID, "the end-of-block symbol (256) needs a code");
    return install_tree(ctx, lit_len, dist_len);
}

int hdlz_tree_lengths(const uint64_t *count, int n, int max_bits, uint8_t *len)
{
    if (!count || !len || n < 1 || n > 288 || max_bits < 1 || max_bits > 15)
        return set_error(HDLZ_ERR_INVALID, "hdlz_tree_lengths: 1..288 symbols, 1..15 bits");
    int used = 0;
    for (int s = 0; s < n; s++) used += count[s] != 0;
    if (used > (1 << max_bits)) return set_error(HDLZ_ERR_INVALID, "%d symbols do not fit %d bits", used, max_bits);
    limited_lengths(count, n, max_bits, len);
    return HDLZ_SUCCESS;
}

int hdlz_tree_header(const uint8_t *lit_len, const uint8_t *dist_len, int container, uint8_t *out, uint32_t out_cap,
                     uint32_t *out_bits)
{
    if (!lit_len || !dist_len || !out_bits) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (container < HDLZ_CONTAINER_ZLIB || container > HDLZ_CONTAINER_GZIP)
        return set_error(HDLZ_ERR_INVALID, "unknown container %d", container);
    int rc;
    if ((rc = check_code(lit_len, 286, "literal/length"))) return rc;
    if ((rc = check_code(dist_len, 30, "distance"))) return rc;
    BitString bs;
    build_prefix((uint32_t)container, lit_len, dist_len, bs);
    *out_bits = bs.bits;
    if ((bs.bits + 7) / 8 > out_cap) return set_error(HDLZ_ERR_INVALID, "header needs %u bytes", (bs.bits + 7) / 8);
    for (uint32_t i = 0; i < (bs.bits + 7) / 8; i++) out[i] = (uint8_t)(bs.words[i >> 2] >> (8 * (i & 3)));
    return HDLZ_SUCCESS;
}

int hdlz_get_tree(hdlz_ctx *ctx, uint8_t *lit_len, uint8_t *dist_len)
{
    if (!ctx || !ctx->tree_set) return 0;
    if (lit_len) memcpy(lit_len, ctx->tree_lit, 286);
    if (dist_len) memcpy(dist_len, ctx->tree_dist, 30);
    return 1;
}

uint32_t hdlz_compress_bound_tree(hdlz_ctx *ctx, uint32_t len)
{
    if (!ctx) return 0;
    if (!ctx->tree_set) return compress_bound(len, ctx->container);
    DeviceGuard guard;
    if (guard.enter(ctx->device) != cudaSuccess || refresh_tree(ctx)) return 0;
    return tree_bound(ctx, len);
}

int hdlz_train_tree(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                    uint32_t uniform_len, uint64_t n, void *stream)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    DeviceGuard guard;
    HDLZ_CUDA(guard.enter(ctx->device));
    if (n == 0 || !d_in) return set_error(HDLZ_ERR_INVALID, "nothing to train on");
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (in_stride & 15))
        return set_error(HDLZ_ERR_INVALID, "d_in must be 16-byte aligned and in_stride a multiple of 16");
    if (ctx->window != HDLZ_CWINDOW) return set_error(HDLZ_ERR_INVALID, "hdlz_train_tree counts with the FAST (CWINDOW = 32) parse");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *d_hist = nullptr;
    HDLZ_CUDA(cudaMalloc((void **)&d_hist, kTreeHistWords * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_hist, 0, kTreeHistWords * sizeof(unsigned long long), s);
    int rc = e == cudaSuccess ? launch_compress_hist(ctx, d_in, in_stride, d_in_len, uniform_len, n, d_hist, s)
                              : cuda_fail(e, "cudaMemsetAsync");
    unsigned long long h[kTreeHistWords];
    if (!rc) {
        e = cudaMemcpyAsync(h, d_hist, sizeof h, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(e, "histogram read-back");
    }
    cudaFree(d_hist);
    if (rc) return rc;
    // histogram layout: [0..255] literals, [256..287] matches by mask bit f (distance 32 - f), [288..295] by length 3 + n.
    // Every symbol the parse can produce keeps a code (count + 1): a later batch may hold bytes this one did not.
    uint64_t lf[286] = {0}, df[30] = {0};
    for (int b = 0; b < 256; b++) lf[b] = h[b] + 1;
    lf[256] = n + 1;
    for (int k = 0; k < 8; k++) lf[257 + k] = h[kTreeLenBase + k] + 1;
    for (int d = 1; d <= 32; d++) {
        int c = 0;
        while (kDistBase[c + 1] <= d) c++;
        df[c] += h[256 + (32 - d)];
    }
    for (int c = 0; c < 10; c++) df[c] += 1;
    uint8_t lit[286], dist[30];
    limited_lengths(lf, 286, 15, lit);
    limited_lengths(df, 30, 15, dist);
    return install_tree(ctx, lit, dist);
}

int hdlz_compress_stream_dyn(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                             uint32_t *out_len, uint32_t *status)
{
    if (!ctx) return set_error(HDLZ_ERR_INVALID, "null context");
    DeviceGuard guard;
    HDLZ_CUDA(guard.enter(ctx->device));
    if (!in || !out || !out_len) return set_error(HDLZ_ERR_INVALID, "null buffer");
    if (len >= (1u << HDLZ_LMAX)) return set_error(HDLZ_ERR_INVALID, "stream longer than 2^LMAX");
    if (ctx->window != HDLZ_CWINDOW) return set_error(HDLZ_ERR_INVALID, "a stream's own tree needs the FAST compressor (CWINDOW = 32)");
    *out_len = 0;
    // the context's own setting comes back afterwards
    const bool had_tree = ctx->tree_set;
    uint8_t old_lit[286], old_dist[30];
    memcpy(old_lit, ctx->tree_lit, 286);
    memcpy(old_dist, ctx->tree_dist, 30);
    cudaStream_t s = ctx->stream;
    const size_t in_slot = ((size_t)len + 15) & ~(size_t)15;
    int rc;
    if ((rc = grow_device((void **)&ctx->d_in, &ctx->d_in_cap, in_slot + 16))) return rc;
    HDLZ_CUDA(cudaMemcpyAsync(ctx->d_in, in, len, cudaMemcpyHostToDevice, s));
    // the stream's statistics: the parse of its 2 KiB blocks (a block's first position and the window at its start
    // differ from the stream's own parse — a few symbols in 2048 — and every symbol keeps a code anyway)
    const uint64_t nblk = len / 2048;
    if (nblk) rc = hdlz_train_tree(ctx, ctx->d_in, 2048, nullptr, 2048, nblk, s);
    else ctx->tree_set = false;
    if (!rc) {
        const size_t out_slot = ctx->tree_set ? tree_bound(ctx, len) : compress_bound(len, ctx->container);
        rc = grow_device((void **)&ctx->d_out, &ctx->d_out_cap, out_slot);
        if (!rc) rc = grow_device((void **)&ctx->d_meta, &ctx->d_meta_cap, 3 * sizeof(uint32_t));
        if (!rc) rc = hdlz_compress_batch(ctx, ctx->d_in, in_slot ? in_slot : 16, nullptr, len, ctx->d_out, out_slot,
                                          ctx->d_meta + 1, ctx->d_meta + 2, 1, s);
        uint32_t meta[2] = {0, 0};
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(meta, ctx->d_meta + 1, sizeof meta, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) rc = cuda_fail(e, "result read-back");
        }
        if (!rc) {
            uint32_t st = meta[1];
            if (st == HDLZ_OK && meta[0] > out_cap) st = HDLZ_ST_OUT_OVERFLOW;
            if (st == HDLZ_OK) {
                cudaError_t e = cudaMemcpyAsync(out, ctx->d_out, meta[0], cudaMemcpyDeviceToHost, s);
                if (e == cudaSuccess) e = cudaStreamSynchronize(s);
                if (e != cudaSuccess) rc = cuda_fail(e, "stream read-back");
                else *out_len = meta[0];
            }
            if (status) *status = st;
        }
    }
    const int rc2 = had_tree ? install_tree(ctx, old_lit, old_dist) : (ctx->tree_set = false, HDLZ_SUCCESS);
    return rc ? rc : rc2;
}

}  // extern "C"
