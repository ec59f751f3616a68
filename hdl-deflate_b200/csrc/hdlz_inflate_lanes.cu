// hdlz_inflate_lanes.cu — lane-per-stream inflater for batches of small streams made of fixed-
// Huffman and stored blocks (what the reference's own compressor emits: one BTYPE=01 block per
// stream, deflate.py:746-814; and zlib Z_FIXED streams, BASELINE config 3).
//
// Decode states NEXT / INFLATE / COPY of the reference (deflate.py:1402-1445, 1519-1659) with the
// fixed tree (STATIC, :1064-1076; the `stat_leaves` LUT of :151-216 regenerated from RFC 1951):
//   - one THREAD per stream, 128 streams per CTA: symbol decode is inherently serial per stream,
//     so the parallelism of a 2^20-stream batch is across streams, not inside one;
//   - the two fixed-tree LUTs live once per CTA in shared memory;
//   - the warp runs in lock step, one symbol per lane per trip; literals and matches share ONE
//     converged "append up to 4 bytes" step, only the match decode and copies longer than 4 bytes
//     diverge;
//   - produced bytes are packed into 32-bit words held in a 16-word per-lane ring in shared memory
//     (word-interleaved by lane => bank == lane, conflict-free) that serves back-references up to
//     120 bytes — every match of the reference format (CWINDOW = 32) — and is flushed to HBM 32 bytes
//     at a time (two 128-bit stores = whole sectors).  Longer distances read the flushed words back
//     from global memory.
//   - a stream that turns out to need the general decoder (dynamic block, unaligned buffers) is
//     appended to a device work list that the warp-per-stream kernel (hdlz_inflate.cu) consumes.
//
// Algorithmic HBM traffic per stream: C bytes read + L bytes written.

#include "hdlz_common.cuh"

namespace hdlz {

int launch_inflate_general(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                           const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                           uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags,
                           const uint32_t *work_list, const uint32_t *work_count, cudaStream_t s);

namespace {

constexpr int kLWarps = 4;
constexpr int kRing = 32;   // words per lane (31 usable: the slot after the partial word is scratch)

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                      43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_lextra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                     3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                      257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                      12289, 16385, 24577};

// litlen entry: bits 0..3 code length, 4..7 extra bits, 8..9 kind (0 literal, 1 end of block,
// 2 length, 3 invalid), 16.. value (byte or base length)
__device__ uint32_t fixed_lit_entry(uint32_t idx9)
{
    const uint32_t r9 = __brev(idx9) >> 23;          // the 9 bits MSB-first, as RFC 1951 3.2.6 lists codes
    const uint32_t c7 = r9 >> 2, c8 = r9 >> 1;
    uint32_t sym, nb;
    if (c7 <= 0x17) { sym = 256 + c7; nb = 7; }
    else if (c8 <= 0xBF) { sym = c8 - 0x30; nb = 8; }
    else if (c8 <= 0xC7) { sym = 280 + (c8 - 0xC0); nb = 8; }
    else { sym = 144 + (r9 - 0x190); nb = 9; }
    if (sym < 256) return nb | (sym << 16);
    if (sym == 256) return nb | (1u << 8);
    if (sym > 285) return nb | (3u << 8);
    return nb | ((uint32_t)c_lextra[sym - 257] << 4) | (2u << 8) | ((uint32_t)c_lbase[sym - 257] << 16);
}

// distance entry (index = the 5 code bits as they sit in the stream): bits 0..3 extra bits,
// 8.. base distance; 0xF in the low nibble marks the invalid symbols 30 and 31
__device__ uint32_t fixed_dist_entry(uint32_t idx5)
{
    const uint32_t d = __brev(idx5) >> 27;
    if (d >= 30) return 0xFu;
    return (d < 2 ? 0u : (d >> 1) - 1u) | ((uint32_t)c_dbase[d] << 8);
}

__global__ void __launch_bounds__(kLWarps * 32)
k_inflate_lanes(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint64_t in_stride,
                const uint32_t *__restrict__ in_len, uint8_t *out, uint64_t out_stride, uint32_t out_cap,
                uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams, uint32_t flags,
                uint32_t *__restrict__ work_list, uint32_t *__restrict__ work_count)
{
    __shared__ uint32_t s_lit[512];
    __shared__ uint32_t s_dist[32];
    __shared__ uint32_t s_ring[kLWarps][kRing][32];

    for (int i = threadIdx.x; i < 512; i += kLWarps * 32) s_lit[i] = fixed_lit_entry((uint32_t)i);
    if (threadIdx.x < 32) s_dist[threadIdx.x] = fixed_dist_entry(threadIdx.x);
    __syncthreads();

    // per-lane decoder state.  The warp iterates in lock step: every trip of the main loop each live
    // lane decodes ONE symbol (and performs its copy); the vote that controls the loop is the
    // reconvergence point, so lanes that took the literal and the match branch meet again each trip.
    enum { S_HEADER = 0, S_FIXED = 1, S_STORED = 2, S_DONE = 3 };

    const uint64_t sid = (uint64_t)blockIdx.x * (kLWarps * 32) + threadIdx.x;
    const bool valid = sid < n_streams;
    uint32_t *ring = &s_ring[threadIdx.x >> 5][0][threadIdx.x & 31];     // word k at ring[k * 32]

    const uint32_t n_in = valid ? in_len[sid] : 0;
    const uint8_t *src = in + (valid ? (in_off ? in_off[sid] : sid * in_stride) : 0);
    uint8_t *dst = out + (valid ? sid * out_stride : 0);
    uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst);
    const uint32_t *inw = reinterpret_cast<const uint32_t *>(src);

    uint32_t st = HDLZ_OK;
    bool hand_over = false;
    uint32_t o = 0, cw = 0, flushed = 0;
    uint32_t ad_a = 1, ad_b = 0, ad_n = 0;
    const bool want_adler = (flags & HDLZ_F_VERIFY_ADLER) != 0;
    uint32_t state = S_HEADER;

    if (!valid) {
        state = S_DONE;
    } else if ((reinterpret_cast<uintptr_t>(src) & 3u) || (reinterpret_cast<uintptr_t>(dst) & 15u)) {
        hand_over = true;                              // the warp-per-stream kernel takes any alignment
        state = S_DONE;
    } else if (n_in < 2) {
        st = HDLZ_ST_TRUNCATED;
        state = S_DONE;
    } else if (flags & HDLZ_F_VERIFY_HEADER) {
        const uint32_t cmf = src[0], flg = src[1];
        if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u || (flg & 0x20u)) {
            st = HDLZ_ST_BAD_HEADER;
            state = S_DONE;
        }
    }

    const uint32_t nfull = n_in >> 2;
    uint32_t wi = 1;
    uint64_t acc = 0;
    uint32_t fill = 16;
    uint32_t final_blk = 0;
    uint32_t stored_left = 0;

    auto load_word = [&](uint32_t w) -> uint32_t {
        if (w < nfull) return __ldg(inw + w);
        uint32_t v = 0;
        if (w == nfull)
            for (uint32_t b = 0; b < (n_in & 3u); ++b) v |= (uint32_t)src[4 * w + b] << (8 * b);
        return v;
    };
    auto refill = [&]() {          // call when fill < 32
        acc |= (uint64_t)load_word(wi) << fill;
        ++wi;
        fill += 32;
    };
    auto adler_bytes = [&](uint32_t v, uint32_t m) {
        for (uint32_t k = 0; k < m; ++k) {
            ad_a += (v >> (8u * k)) & 255u;
            ad_b += ad_a;
            if (++ad_n == 5552) { ad_a %= 65521u; ad_b %= 65521u; ad_n = 0; }
        }
    };
    // Output word x: the ring always holds the last kRing words INCLUDING the partial word being
    // filled, so recent data needs no special case; older words come back from global memory
    // (they were flushed: a 32-byte group is stored as soon as its last word completes).
    auto fetch = [&](uint32_t x, uint32_t wo) -> uint32_t {
        if (wo - x < (uint32_t)(kRing - 1)) return ring[(x & (kRing - 1)) * 32];
        return dst32[x];
    };
    // Store 32 completed bytes (two 128-bit stores = whole sectors) if this lane has them.
    // Called at converged points of the loop; `flushed` counts the words already in global memory.
    auto flush8 = [&]() {
        while ((o >> 2) - flushed >= 8u) {
            uint32_t q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = ring[((flushed + k) & (kRing - 1)) * 32];
            uint4 *g = reinterpret_cast<uint4 *>(dst32 + flushed);
            g[0] = make_uint4(q[0], q[1], q[2], q[3]);
            g[1] = make_uint4(q[4], q[5], q[6], q[7]);
            flushed += 8u;
        }
    };
    // append the low m (1..4) bytes of v to the output; branch-free: the word being filled and
    // the (possibly empty) next one are both written back to the ring
    auto append = [&](uint32_t v, uint32_t m) {
        v &= 0xFFFFFFFFu >> (32u - 8u * m);
        if (want_adler) adler_bytes(v, m);
        const uint32_t ob = o & 3u, wo = o >> 2;
        const uint64_t comb = (uint64_t)cw | ((uint64_t)v << (8u * ob));
        o += m;
        ring[(wo & (kRing - 1)) * 32] = (uint32_t)comb;
        ring[((wo + 1u) & (kRing - 1)) * 32] = (uint32_t)(comb >> 32);
        cw = ob + m >= 4u ? (uint32_t)(comb >> 32) : (uint32_t)comb;
    };
    // bytes s .. s+3 of the output for a back-reference of distance `dist` (s = o - dist)
    auto source = [&](uint32_t dist) -> uint32_t {
        const uint32_t s = o - dist;
        const uint32_t ws = s >> 2, wo = o >> 2;
        uint32_t w0, w1;
        if (wo - ws < (uint32_t)(kRing - 1)) {         // both words are in the ring (ws + 1 <= wo + 1: scratch slot)
            w0 = ring[(ws & (kRing - 1)) * 32];
            w1 = ring[((ws + 1u) & (kRing - 1)) * 32];
        } else {                                       // far back-reference: flushed long ago
            w0 = dst32[ws];
            w1 = dst32[ws + 1u];
        }
        const uint32_t v = __funnelshift_r(w0, w1, 8u * (s & 3u));
        // distance < 4: the source overlaps what is being written -> period-`dist` pattern:
        // keep the first `dist` bytes and replicate them with one multiply
        const uint32_t dd = dist < 4u ? dist : 4u;
        const uint32_t keep = 0xFFFFFFFFu >> ((32u - 8u * dd) & 31u);   // dd == 0 only on lanes that are not copying
        const uint32_t mult = dd == 4u ? 1u : dd == 3u ? 0x01000001u : dd == 2u ? 0x00010001u : 0x01010101u;
        return (v & keep) * mult;
    };

    if (state != S_DONE) {
        acc = (uint64_t)(load_word(0) >> 16);          // skip the zlib header: di = 2 (deflate.py:644)
        ring[0] = 0;
    }

    uint32_t rem = 0, dist = 1;      // bytes still to copy of the current match, its distance
    uint32_t trip = 0;

    while (__any_sync(HDLZ_FULL_MASK, state != S_DONE)) {
        // every 8 trips (a trip appends at most 4 bytes, so at most 8 words accumulate) all lanes
        // store their completed 32-byte groups together
        if ((++trip & 7u) == 0) flush8();
        if (state == S_FIXED) {
            // ---- fixed block (NEXT / INFLATE / COPY): every trip appends at most four bytes ----
            // A lane either continues the copy it is in (rem != 0) or decodes: up to four
            // consecutive literals (packed into one append), or one match / end-of-block code.
            if (fill < 32) {
                if (wi > nfull + 2) { st = HDLZ_ST_TRUNCATED; state = S_DONE; }
                else refill();
            }
            const bool decode = state == S_FIXED && rem == 0;
            const uint32_t room = out_cap - o;                       // o <= out_cap always
            // literal run: symbol j is taken while everything before it was a literal, its code
            // fits in the 32 valid bits, and the output has room
            uint32_t used = 0, lits = 0, nlit = 0;
            const uint32_t a32 = (uint32_t)acc;                      // >= 32 valid bits
            const uint32_t e0 = s_lit[a32 & 511u];
            {
                bool go = decode;
                uint32_t e = e0, x = a32;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k) e = s_lit[x & 511u];
                    const uint32_t nb = e & 15u;
                    go = go && (e & 0x300u) == 0 && used + nb <= 32u && nlit < room;
                    lits |= go ? ((e >> 16) << (8 * k)) : 0u;
                    used += go ? nb : 0u;
                    nlit += go ? 1u : 0u;
                    x >>= nb;                                        // only meaningful while go holds
                }
            }
            if (decode && nlit == 0) {
                // first symbol is not a literal (or no room): <length code><extra><5-bit distance code><extra>,
                // at most 8 + 5 + 5 + 13 = 31 bits, all inside the low word of the bit buffer
                const uint32_t nb = e0 & 15u, eb = (e0 >> 4) & 15u, kind = (e0 >> 8) & 3u, base = e0 >> 16;
                const uint32_t x1 = a32 >> nb;
                const uint32_t len = base + (x1 & ((1u << eb) - 1u));
                const uint32_t x2 = x1 >> eb;
                const uint32_t de = s_dist[x2 & 31u];
                const uint32_t deb = de & 15u;
                const uint32_t dnew = (de >> 8) + ((x2 >> 5) & ((1u << deb) - 1u));
                if (kind == 2u) {
                    used = nb + eb + 5u + deb;
                    if (deb == 15u) { st = HDLZ_ST_BAD_CODE; state = S_DONE; }
                    else if (dnew > o) { st = HDLZ_ST_DIST_TOO_FAR; state = S_DONE; }     // "distance too big" (deflate.py:1506-1508)
                    else if (len > room) { st = HDLZ_ST_OUT_OVERFLOW; state = S_DONE; }
                    else { rem = len; dist = dnew; }
                } else if (kind == 1u) {
                    used = nb;
                    state = final_blk ? S_DONE : S_HEADER;          // end of block
                    if (final_blk) final_blk = 2;                   // 2 = finished cleanly
                } else if (kind == 0u) {
                    st = HDLZ_ST_OUT_OVERFLOW;                      // a literal with no room left
                    state = S_DONE;
                } else {
                    st = HDLZ_ST_BAD_CODE;                          // "invalid token" (deflate.py:1559-1560)
                    state = S_DONE;
                }
            }
            acc >>= used; fill -= used;
            // one append step: the literals, or up to four bytes of the copy (COPY, deflate.py:1627-1656)
            const bool copying = rem != 0;
            const uint32_t sv = source(copying ? dist : 0u);
            const uint32_t m = copying ? (rem < 4u ? rem : 4u) : nlit;
            rem -= copying ? m : 0u;
            if (m && state != S_DONE) append(copying ? sv : lits, m);
        } else if (state == S_HEADER) {
            if (fill < 32) refill();
            final_blk = (uint32_t)acc & 1u;
            const uint32_t type = ((uint32_t)acc >> 1) & 3u;
            acc >>= 3; fill -= 3;
            if (type == 1) {
                state = S_FIXED;
            } else if (type == 2) {
                // dynamic block: this stream belongs to the general decoder, which restarts it
                // from its first byte (anything produced so far is simply rewritten)
                hand_over = true;
                state = S_DONE;
            } else if (type == 3) {
                st = HDLZ_ST_BAD_BTYPE;                             // "Bad method" (deflate.py:718-721)
                state = S_DONE;
            } else {
                // stored block header (deflate.py:709-717)
                const uint32_t drop = fill & 7u;
                acc >>= drop; fill -= drop;
                if (fill < 32) refill();
                const uint32_t len = (uint32_t)acc & 0xFFFFu, nlen = ((uint32_t)acc >> 16) & 0xFFFFu;
                acc >>= 32; fill -= 32;
                const uint64_t bytepos = ((uint64_t)wi * 32 - fill) >> 3;
                if ((len ^ 0xFFFFu) != nlen) { st = HDLZ_ST_BAD_STORED; state = S_DONE; }
                else if (bytepos + len > n_in) { st = HDLZ_ST_TRUNCATED; state = S_DONE; }
                else if ((uint64_t)o + len > out_cap) { st = HDLZ_ST_OUT_OVERFLOW; state = S_DONE; }
                else { stored_left = len; state = S_STORED; }
            }
        } else if (state == S_STORED) {
            // stored bytes, up to 4 per trip (COPY with method 0, deflate.py:1603-1616)
            for (int k = 0; k < 4 && stored_left; ++k, --stored_left) {
                if (fill < 8) refill();
                append((uint32_t)acc & 255u, 1);
                acc >>= 8; fill -= 8;
            }
            if (stored_left == 0) {
                state = final_blk ? S_DONE : S_HEADER;
                if (final_blk) final_blk = 2;
            }
        }
    }

    if (valid && st == HDLZ_OK && !hand_over) {
        if (final_blk != 2) {
            st = HDLZ_ST_TRUNCATED;
        } else {
            // words completed since the last 32-byte flush, then the bytes of the partial word
            const uint32_t wo = o >> 2;
            for (uint32_t x = flushed; x < wo; ++x) dst32[x] = ring[(x & (kRing - 1)) * 32];
            for (uint32_t k = 0; k < (o & 3u); ++k) dst[(o & ~3u) + k] = (uint8_t)(cw >> (8 * k));
            const uint64_t bp = (uint64_t)wi * 32 - fill;
            const uint64_t tp = (bp + 7) >> 3;                       // Adler-32 trailer must be present
            if (bp > 8ull * n_in || tp + 4 > n_in) {
                st = HDLZ_ST_TRUNCATED;                              // "NO EOF!" (deflate.py:1535-1539)
            } else if (want_adler) {
                ad_a %= 65521u; ad_b %= 65521u;
                const uint32_t want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) |
                                      ((uint32_t)src[tp + 2] << 8) | src[tp + 3];
                if (((ad_b << 16) | ad_a) != want) st = HDLZ_ST_BAD_ADLER;
            }
        }
    }

    if (!valid) return;
    if (hand_over) {
        work_list[atomicAdd(work_count, 1u)] = (uint32_t)sid;
    } else {
        out_len[sid] = st == HDLZ_OK ? o : 0;
        if (status) status[sid] = st;
    }
}

}  // namespace

int launch_inflate(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                   const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                   uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags, uint32_t *d_work,
                   cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    // Few streams: one warp each is the better mapping.  Many streams: one lane each first,
    // the warp-per-stream kernel then finishes whatever was handed over.
    const bool lanes_first = !(flags & HDLZ_F_FORCE_GENERAL) && (n >= 1024 || (flags & HDLZ_F_FORCE_LANES)) &&
                             n < 0xFFFFFFFFull && d_work != nullptr;
    if (!lanes_first)
        return launch_inflate_general(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len,
                                      d_status, n, flags, nullptr, nullptr, s);
    uint32_t *count = d_work, *list = d_work + 8;
    HDLZ_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), s));
    const uint64_t blocks = (n + kLWarps * 32 - 1) / (kLWarps * 32);
    k_inflate_lanes<<<(unsigned)blocks, kLWarps * 32, 0, s>>>(d_in, d_in_off, in_stride, d_in_len, d_out, out_stride,
                                                              out_cap, d_out_len, d_status, n, flags, list, count);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return launch_inflate_general(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len,
                                  d_status, n, flags, list, count, s);
}

}  // namespace hdlz
