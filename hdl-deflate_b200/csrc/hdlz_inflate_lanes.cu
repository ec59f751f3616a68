// hdlz_inflate_lanes.cu — lane-per-stream inflater for batches of many streams: fixed-Huffman and
// stored blocks (what the reference's own compressor emits: one BTYPE=01 block per stream,
// deflate.py:746-814; zlib Z_FIXED streams, BASELINE config 3) and dynamic-Huffman blocks (zlib
// level 6, BASELINE config 4; BL / READBL / REPEAT / HF1..HF4, deflate.py:1084-1400).
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
//   - the dynamic-capable instantiation is persistent: when all lanes of a warp are between streams
//     they take their next 32 streams together;
//   - dynamic blocks: every resident thread owns a 1.8 KiB scratch in global memory (8-bit literal/
//     length and 7-bit distance primary tables + canonical arrays for longer codes; smaller tables
//     keep more of them in L2: 9 -> 8 bits took config 4 from 59 to 68 GB/s, 7 bits lose it to the slow
//     path); the lane parses the block header and builds its tables by itself (plain scalar
//     code, all lanes of a warp usually do it at the same time);
//   - a stream the lanes cannot take (unaligned buffers, no scratch) is appended to a device work
//     list that the warp-per-stream kernel (hdlz_inflate.cu) consumes.
//
// Algorithmic HBM traffic per stream: C bytes read + L bytes written.

#include "hdlz_common.cuh"
#include "hdlz_frame.cuh"
#include "hdlz_split.cuh"

namespace hdlz {

int launch_inflate_general(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                           const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                           uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags,
                           const uint32_t *work_list, const uint32_t *work_count, cudaStream_t s);

namespace {

constexpr int kLWarps = 4;
constexpr int kLaneCtasPerSm = 10;     // fixed/stored instantiation (48 registers)
constexpr int kDynCtasPerSm = 6;       // dynamic-capable instantiation (80 registers, 1.8 KiB of scratch per thread)
constexpr int kRing = 32;   // words per lane (30 usable: the two slots after the partial word are scratch)
constexpr int kDynLitBits = 8;
constexpr int kDynDistBits = 7;

// per resident thread, global memory.  Table entry: (symbol << 4) | code length, 0 = longer code.
// The primary tables every symbol goes through live in their own array (one LaneHot per resident
// thread) so that the launch can ask the L2 to keep exactly that range (access-policy window).
struct LaneHot {
    uint16_t lit[1 << kDynLitBits];
    uint16_t dist[1 << kDynDistBits];
};

struct LaneScratch {
    uint16_t sorted_l[288];
    uint16_t sorted_d[32];
    uint16_t cnt_l[16];
    uint16_t cnt_d[16];
    uint16_t resume_l[2];
    uint16_t resume_d[2];
    uint8_t lens[320];
};

__constant__ uint8_t c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                      43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_lextra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                     3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                      257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                      12289, 16385, 24577};

// fixed litlen entry, indexed by 9 stream bits.
//   literal     : bits 0..3 code length (8 | 9), bit 8 set, 16..23 the byte
//   non-literal : bits 0..3 ZERO (so a literal-decode chain that meets one stops consuming: the
//                 following look-ups see the same bits again), 4..7 code length, 9..10 kind
//                 (1 end of block, 2 length, 3 invalid), 11..14 extra bits, 16.. base length
__device__ uint32_t fixed_lit_entry(uint32_t idx9)
{
    const uint32_t r9 = __brev(idx9) >> 23;          // the 9 bits MSB-first, as RFC 1951 3.2.6 lists codes
    const uint32_t c7 = r9 >> 2, c8 = r9 >> 1;
    uint32_t sym, nb;
    if (c7 <= 0x17) { sym = 256 + c7; nb = 7; }
    else if (c8 <= 0xBF) { sym = c8 - 0x30; nb = 8; }
    else if (c8 <= 0xC7) { sym = 280 + (c8 - 0xC0); nb = 8; }
    else { sym = 144 + (r9 - 0x190); nb = 9; }
    if (sym < 256) return nb | 0x100u | (sym << 16);
    if (sym == 256) return (nb << 4) | (1u << 9);
    if (sym > 285) return (nb << 4) | (3u << 9);
    return (nb << 4) | (2u << 9) | ((uint32_t)c_lextra[sym - 257] << 11) | ((uint32_t)c_lbase[sym - 257] << 16);
}

// distance entry (index = the 5 code bits as they sit in the stream): bits 0..3 extra bits,
// 8.. base distance; 0xF in the low nibble marks the invalid symbols 30 and 31
__device__ uint32_t fixed_dist_entry(uint32_t idx5)
{
    const uint32_t d = __brev(idx5) >> 27;
    if (d >= 30) return 0xFu;
    return (d < 2 ? 0u : (d >> 1) - 1u) | ((uint32_t)c_dbase[d] << 8);
}

// literal/length symbol -> the entry format of fixed_lit_entry without the code length
__device__ uint32_t sym_entry(uint32_t sym)
{
    if (sym < 256) return sym << 16;
    if (sym == 256) return 1u << 8;
    if (sym > 285) return 3u << 8;
    return ((uint32_t)c_lextra[sym - 257] << 4) | (2u << 8) | ((uint32_t)c_lbase[sym - 257] << 16);
}

// Canonical Huffman tables of one code (HF1INIT..HF4, SPREAD; deflate.py:1227-1400), built by ONE
// thread in its scratch.  Returns 0, or 1 for an over-subscribed / illegally incomplete code
// (zlib's inflate_table rules).
__device__ __noinline__ int lane_build(const uint8_t *lens, int nsym, uint16_t *tbl, int tbits, uint16_t *cnt,
                                       uint16_t *sorted, uint16_t *resume, bool allow_incomplete)
{
    uint16_t first[16], offs[16], run[16];
    for (int l = 0; l < 16; ++l) { cnt[l] = 0; run[l] = 0; }
    resume[0] = resume[1] = 0;      // a code without symbols must not leave the previous build's resume point behind
    for (int s = 0; s < nsym; ++s) cnt[lens[s]]++;
    uint32_t *t32 = reinterpret_cast<uint32_t *>(tbl);
    for (int i = 0; i < (1 << tbits) / 2; ++i) t32[i] = 0;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        const int c = cnt[l];
        left = 2 * left - c;
        if (c) maxlen = l;
        if (left < 0) return 1;
    }
    if (maxlen == 0) return 0;                 // no codes: any use fails later
    if (left > 0 && !(allow_incomplete && maxlen == 1)) return 1;
    uint32_t code = 0, off = 0;
    for (int l = 1; l <= 15; ++l) {
        first[l] = (uint16_t)code;
        offs[l] = (uint16_t)off;
        code = (code + cnt[l]) << 1;
        off += cnt[l];
    }
    // where the bit-serial decode of a code longer than the table resumes: its `first` / `index` after
    // the lengths 1 .. tbits have been ruled out
    resume[0] = tbits < 15 ? first[tbits + 1] : 0;
    resume[1] = tbits < 15 ? offs[tbits + 1] : 0;
    for (int s = 0; s < nsym; ++s) {
        const uint32_t l = lens[s];
        if (!l) continue;
        const uint32_t k = run[l]++;
        sorted[offs[l] + k] = (uint16_t)s;
        if ((int)l <= tbits) {
            const uint32_t rev = __brev(first[l] + k) >> (32 - l);
            const uint16_t ent = (uint16_t)((s << 4) | l);
            for (uint32_t idx = rev; idx < (1u << tbits); idx += 1u << l) tbl[idx] = ent;
        }
    }
    return 0;
}

// code longer than the primary table: canonical decode one bit at a time, starting after the `tbits`
// bits the table has already ruled out.  -> (sym << 4) | len, 0 = invalid
__device__ __noinline__ uint32_t lane_slow_decode(uint32_t bits, const uint16_t *cnt, const uint16_t *sorted, int tbits,
                                                  const uint16_t *resume)
{
    int code = (int)((__brev(bits) >> (32 - tbits)) << 1), first = resume[0], index = resume[1];
    for (int l = tbits + 1; l <= 15; ++l) {
        code |= (int)((bits >> (l - 1)) & 1u);
        const int c = cnt[l];
        if (code - c < first) return ((uint32_t)sorted[index + (code - first)] << 4) | (uint32_t)l;
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0;
}

// kDyn = false: fixed / stored blocks only (lean: fewer registers, more resident warps); a stream
//                with a dynamic block is appended to `dyn_list`.  Items: all n_streams.
// kDyn = true : everything; items come from `items[0 .. *item_count)` (the list the first
//                instantiation filled).  Either way unaligned streams go to `work_list`.
template <bool kDyn>
__global__ void __launch_bounds__(kLWarps * 32)
k_inflate_lanes(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint64_t in_stride,
                const uint32_t *__restrict__ in_len, uint8_t *out, uint64_t out_stride, uint32_t out_cap,
                uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams, uint32_t flags,
                uint32_t *__restrict__ work_list, uint32_t *__restrict__ work_count,
                uint32_t *__restrict__ dyn_list, uint32_t *__restrict__ dyn_count,
                const uint32_t *__restrict__ items, const uint32_t *__restrict__ item_count, uint32_t item_skip,
                LaneScratch *scratch, LaneHot *hot_base)
{
    // with a list, the first `item_skip` entries belong to another kernel (the two-phase route)
    const uint64_t n_items = items ? (uint64_t)(*item_count > item_skip ? *item_count - item_skip : 0u) : n_streams;
    if (n_items == 0) return;
    __shared__ uint32_t s_lit[512];       // fixed tree: 9 stream bits -> entry (see fixed_lit_entry)
    __shared__ uint32_t s_dist[32];       // fixed tree: 5 stream bits -> distance entry
    __shared__ uint32_t s_sym[288];       // literal/length symbol -> entry without code length
    __shared__ uint32_t s_dsym[32];       // distance symbol -> distance entry
    __shared__ uint32_t s_nib[16];        // CRC-32 nibble table (gzip trailer check)
    __shared__ uint2 s_pat[8];            // min(distance, 4) -> {bytes of the source word to keep, replication factor}
    __shared__ unsigned long long s_mask8[9];   // n -> mask of the low n bytes
    __shared__ unsigned long long s_mult8[9];   // min(distance, 8) -> factor that repeats the low `distance` bytes over 8
    __shared__ uint32_t s_ring[kLWarps][kRing][32];

    for (int i = threadIdx.x; i < 512; i += kLWarps * 32) s_lit[i] = fixed_lit_entry((uint32_t)i);
    for (int i = threadIdx.x; i < 288; i += kLWarps * 32) s_sym[i] = sym_entry((uint32_t)i);
    if (threadIdx.x < 32) {
        s_dist[threadIdx.x] = fixed_dist_entry(threadIdx.x);
        s_dsym[threadIdx.x] = fixed_dist_entry(__brev(threadIdx.x) >> 27);
    }
    if (threadIdx.x < 16) s_nib[threadIdx.x] = crc32_nibble_entry(threadIdx.x);
    if (threadIdx.x < 9) {
        const uint32_t d = threadIdx.x;
        s_mask8[d] = d == 0 ? 0ull : d == 8 ? ~0ull : (1ull << (8 * d)) - 1ull;
        unsigned long long mul = 0;
        for (uint32_t sh = 0; d && sh < 64; sh += 8 * d) mul |= 1ull << sh;
        s_mult8[d] = mul;
    }
    if (threadIdx.x < 8) {
        const uint32_t d = threadIdx.x;
        s_pat[d] = d == 1 ? make_uint2(0xFFu, 0x01010101u) : d == 2 ? make_uint2(0xFFFFu, 0x00010001u)
                   : d == 3 ? make_uint2(0xFFFFFFu, 0x01000001u) : d == 4 ? make_uint2(0xFFFFFFFFu, 1u)
                                                                          : make_uint2(0u, 0u);
    }
    __syncthreads();

    // per-lane decoder state machine.  The warp iterates in lock step: the vote that controls the
    // loop is the reconvergence point, so lanes that took different branches meet again each trip.
    enum { S_IDLE = 0, S_HEADER = 1, S_FIXED = 2, S_STORED = 3, S_DYN = 4, S_FINISH = 5, S_DONE = 6 };

    const uint64_t n_threads = (uint64_t)gridDim.x * (kLWarps * 32);
    const uint64_t gtid = (uint64_t)blockIdx.x * (kLWarps * 32) + threadIdx.x;
    uint64_t next_item = gtid;
    uint64_t sid = 0;
    uint32_t *ring = &s_ring[threadIdx.x >> 5][0][threadIdx.x & 31];     // word k at ring[k * 32]
    LaneScratch *my = (kDyn && scratch) ? scratch + gtid : nullptr;
    LaneHot *hot = (kDyn && scratch) ? hot_base + gtid : nullptr;
    bool to_dyn = false;            // hand this stream to the dynamic-capable instantiation

    // the container checksum, only on request: Adler-32 rides along with the appends, the CRC-32 of a gzip
    // member is taken over the finished output
    const bool want_adler = (flags & HDLZ_F_VERIFY_ADLER) && !(flags & (HDLZ_F_RAW | HDLZ_F_GZIP));
    const bool want_crc = (flags & HDLZ_F_VERIFY_ADLER) && (flags & HDLZ_F_GZIP);
    const uint32_t trailer_bytes = (flags & HDLZ_F_RAW) ? 0u : (flags & HDLZ_F_GZIP) ? 8u : 4u;
    uint32_t state = S_IDLE;

    // per-stream state
    uint32_t n_in = 0, nfull = 0, tailw = 0, nextw = 0;
    const uint8_t *src = in;
    uint8_t *dst = out;
    uint32_t *dst32 = reinterpret_cast<uint32_t *>(out);
    const uint32_t *inw = reinterpret_cast<const uint32_t *>(in);
    uint32_t st = HDLZ_OK;
    bool hand_over = false;
    uint32_t o = 0, cw = 0, flushed = 0;
    uint32_t ad_a = 1, ad_b = 0, ad_n = 0;
    uint32_t wi = 1, fill = 16, final_blk = 0, stored_left = 0;
    uint64_t acc = 0;
    uint32_t rem = 0, dist = 1;      // bytes still to copy of the current match, its distance
    uint32_t trip = 0;

    // input word w of the stream; the last, partial word was assembled once when the stream was opened
    auto load_word = [&](uint32_t w) -> uint32_t {
        uint32_t v = w == nfull ? tailw : 0u;
        if (w < nfull) v = __ldg(inw + w);
        return v;
    };
    auto refill = [&]() {          // call when fill < 32.  `nextw` is word `wi`, fetched one refill ahead
        acc |= (uint64_t)nextw << fill;
        ++wi;
        fill += 32;
        nextw = load_word(wi);
    };
    // Adler-32 of m appended bytes (the low m bytes of v, the rest zero): a += S, b += m a + m S - sum j x_j with
    // two dot products per word; the modulo is deferred like zlib's (both sums stay below 2^32 for 5552 bytes)
    auto adler_bytes = [&](uint32_t v, uint32_t m) {
        const uint32_t S = __dp4a(v, 0x01010101u, 0u), J = __dp4a(v, 0x03020100u, 0u);
        ad_b += m * (ad_a + S) - J;
        ad_a += S;
        ad_n += m;
        if (ad_n >= 5544u) { ad_a %= 65521u; ad_b %= 65521u; ad_n = 0; }
    };
    auto adler_bytes8 = [&](unsigned long long v, uint32_t m) {
        const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        const uint32_t S = __dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
        const uint32_t J = __dp4a(lo, 0x03020100u, __dp4a(hi, 0x07060504u, 0u));
        ad_b += m * (ad_a + S) - J;
        ad_a += S;
        ad_n += m;
        if (ad_n >= 5544u) { ad_a %= 65521u; ad_b %= 65521u; ad_n = 0; }
    };
    // Store 32 completed bytes (two 128-bit stores = whole sectors) while this lane has them.
    // Called at converged points of the loop; `flushed` counts the words already in global memory.
    auto flush8 = [&]() {
        while ((o >> 2) - flushed >= 8u) {
            uint32_t q[8];
            const uint32_t *g8 = ring + (flushed & (kRing - 1)) * 32;     // `flushed` is a multiple of 8: no wrap inside a group
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = g8[k * 32];
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
    // bytes s .. s+3 of the output for a back-reference of distance `d` (s = o - d).  The ring always
    // holds the last kRing - 1 words including the partial one; older words were flushed and come
    // back from global memory.
    auto source = [&](uint32_t d) -> uint32_t {
        const uint32_t s = o - d;
        const uint32_t ws = s >> 2, wo = o >> 2;
        uint32_t w0, w1;
        if (wo - ws < (uint32_t)(kRing - 2)) {          // the eight-byte append uses two slots past the partial word
            w0 = ring[(ws & (kRing - 1)) * 32];
            w1 = ring[((ws + 1u) & (kRing - 1)) * 32];
        } else {
            w0 = dst32[ws];
            w1 = dst32[ws + 1u];
        }
        const uint32_t v = __funnelshift_r(w0, w1, 8u * (s & 3u));
        // distance < 4: the source overlaps what is being written -> period-`d` pattern:
        // keep the first `d` bytes and replicate them with one multiply (factors from a small table;
        // d == 0 only on lanes that are not copying)
        const uint2 pat = s_pat[d < 4u ? d : 4u];
        return (v & pat.x) * pat.y;
    };
    // eight-byte forms for the fixed-block trip: bytes s .. s+7 of the output (s = o - d) ...
    auto source8 = [&](uint32_t d) -> unsigned long long {
        const uint32_t s = o - d;
        const uint32_t ws = s >> 2, wo = o >> 2;
        uint32_t w0, w1, w2;
        if (wo - ws < (uint32_t)(kRing - 2)) {
            w0 = ring[(ws & (kRing - 1)) * 32];
            w1 = ring[((ws + 1u) & (kRing - 1)) * 32];
            w2 = ring[((ws + 2u) & (kRing - 1)) * 32];
        } else {
            w0 = dst32[ws];
            w1 = dst32[ws + 1u];
            w2 = dst32[ws + 2u];
        }
        const uint32_t sh = 8u * (s & 3u);
        const unsigned long long v = (unsigned long long)__funnelshift_r(w0, w1, sh) |
                                     ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
        // distance < 8: keep the first `d` bytes and repeat them (d == 0 only on lanes that are not copying)
        const uint32_t dd = d < 8u ? d : 8u;
        return (v & s_mask8[dd]) * s_mult8[dd];
    };
    // ... and appending the low m (1..8) bytes of v: three ring words are written back
    auto append8 = [&](unsigned long long v, uint32_t m) {
        v &= s_mask8[m];
        if (want_adler) adler_bytes8(v, m);
        const uint32_t ob = o & 3u, wo = o >> 2, sh = 8u * ob;
        const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        const uint32_t wa = cw | (lo << sh);
        const uint32_t wb = __funnelshift_l(lo, hi, sh);
        const uint32_t wc = __funnelshift_l(hi, 0u, sh);
        o += m;
        ring[(wo & (kRing - 1)) * 32] = wa;
        ring[((wo + 1u) & (kRing - 1)) * 32] = wb;
        ring[((wo + 2u) & (kRing - 1)) * 32] = wc;
        const uint32_t nw = (ob + m) >> 2;
        cw = nw == 0u ? wa : nw == 1u ? wb : wc;
    };
    auto fail = [&](uint32_t code) { st = code; state = S_FINISH; };

    while (__any_sync(HDLZ_FULL_MASK, state != S_DONE)) {
        // every 8 trips (a trip appends at most 8 bytes, so at most 16 words accumulate on top of 7) all lanes
        // store their completed 32-byte groups together
        if ((++trip & 7u) == 0) flush8();

        // ---- when every lane of the warp is between streams, all take their next stream together
        // (a per-lane refetch would run this and the block header with one or two active lanes)
        // (looked at every eighth trip: the vote is not free and a finished lane can wait that long)
        if ((trip & 7u) == 1u && __all_sync(HDLZ_FULL_MASK, state == S_IDLE || state == S_DONE)) {
            if (state == S_IDLE) {
            if (next_item >= n_items) {
                state = S_DONE;
            } else {
                sid = items ? (uint64_t)items[next_item + item_skip] : next_item;
                next_item += n_threads;
                to_dyn = false;
                n_in = in_len[sid];
                src = in + (in_off ? in_off[sid] : sid * in_stride);
                dst = out + sid * out_stride;
                dst32 = reinterpret_cast<uint32_t *>(dst);
                inw = reinterpret_cast<const uint32_t *>(src);
                nfull = n_in >> 2;
                st = HDLZ_OK;
                hand_over = false;
                o = 0; cw = 0; flushed = 0;
                ad_a = 1; ad_b = 0; ad_n = 0;
                wi = 1; fill = 16; final_blk = 0; stored_left = 0; rem = 0; dist = 1;
                state = S_HEADER;
                if ((reinterpret_cast<uintptr_t>(src) & 3u) || (reinterpret_cast<uintptr_t>(dst) & 15u)) {
                    hand_over = true;                  // the warp-per-stream kernel takes any alignment
                    state = S_FINISH;
                } else {
                    // container header: zlib (skipped like the reference does: di = 2, deflate.py:644),
                    // none, or gzip
                    const Frame frame = parse_frame(src, n_in, flags);
                    if (frame.status != HDLZ_OK) {
                        fail(frame.status);
                    } else {
                        tailw = 0;
                        for (uint32_t b = 0; b < (n_in & 3u); ++b) tailw |= (uint32_t)src[4 * nfull + b] << (8 * b);
                        const uint32_t skip = 8u * (frame.body & 3u);
                        wi = frame.body >> 2;
                        acc = (uint64_t)(load_word(wi) >> skip);
                        fill = 32u - skip;
                        ++wi;
                        nextw = load_word(wi);
                        ring[0] = 0;
                    }
                }
            }
            }
        }

        if (state == S_FIXED) {
            // ---- fixed block (NEXT / INFLATE / COPY): every trip appends at most eight bytes ----
            // A lane either continues the copy it is in (rem != 0) or decodes: up to four
            // consecutive literals (packed into one append), or one match / end-of-block code.
            // Past the end of the input the bit buffer is fed zeros, which the fixed tree reads as
            // end-of-block: a truncated stream reaches S_FINISH / S_HEADER, where the position is checked.
            if (fill < 32) refill();
            const bool decode = state == S_FIXED && rem == 0;
            const uint32_t room = out_cap - o;                       // o <= out_cap always
            uint32_t used = 0, lits = 0, nlit = 0;
            const uint32_t a32 = (uint32_t)acc;                      // >= 32 valid bits
            const uint32_t e0 = s_lit[a32 & 511u];
            {
                // up to four literals.  No predicates in the chain: a non-literal entry has a zero in
                // the consumed-bits field and a zero byte, so once one is met the remaining look-ups
                // stay on it.  `tally` sums the entries' low fields: bits 0..5 consumed bits, 8..10
                // number of literals.
                uint32_t x = a32, e = e0, tally = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k) e = s_lit[x & 511u];
                    if (k == 3 && (tally & 63u) + (e & 15u) > 32u) e = 0;        // would run past the 32 valid bits
                    lits = __byte_perm(lits, e, 0x3210 ^ ((0x6 ^ k) << (4 * k)));  // byte k of lits = byte 2 of e
                    tally += e & 0x10Fu;
                    x >>= e & 15u;
                }
                if (!decode) tally = 0;                              // the lane is in a copy: nothing is consumed
                used = tally & 63u;
                nlit = tally >> 8;
                if (nlit > room) {
                    // rare: the output ends inside this run of literals
                    used = 0; nlit = 0; lits = 0;
                    x = a32;
                    while (nlit < room) {
                        e = s_lit[x & 511u];
                        if (!(e & 15u)) break;
                        lits |= ((e >> 16) & 255u) << (8u * nlit);
                        used += e & 15u;
                        x >>= e & 15u;
                        ++nlit;
                    }
                }
            }
            if (decode && nlit == 0) {
                // first symbol is not a literal (or no room): <length code><extra><5-bit distance code><extra>,
                // at most 8 + 5 + 5 + 13 = 31 bits, all inside the low word of the bit buffer
                const uint32_t nb = (e0 >> 4) & 15u, eb = (e0 >> 11) & 15u, kind = (e0 >> 9) & 3u, base = e0 >> 16;
                const uint32_t x1 = a32 >> nb;
                const uint32_t len = base + (x1 & ((1u << eb) - 1u));
                const uint32_t x2 = x1 >> eb;
                const uint32_t de = s_dist[x2 & 31u];
                const uint32_t deb = de & 15u;
                const uint32_t dnew = (de >> 8) + ((x2 >> 5) & ((1u << deb) - 1u));
                if (kind == 2u) {
                    used = nb + eb + 5u + deb;
                    if (deb == 15u || dnew > o || len > room) {
                        if (deb == 15u) fail(HDLZ_ST_BAD_CODE);
                        else if (dnew > o) fail(HDLZ_ST_DIST_TOO_FAR);         // "distance too big" (deflate.py:1506-1508)
                        else fail(HDLZ_ST_OUT_OVERFLOW);
                    } else { rem = len; dist = dnew; }
                } else if (kind == 1u) {
                    used = nb;
                    state = final_blk ? S_FINISH : S_HEADER;                // end of block
                    if (final_blk) final_blk = 2;                           // 2 = finished cleanly
                } else if (kind == 0u) {
                    fail(HDLZ_ST_OUT_OVERFLOW);                             // a literal with no room left
                } else {
                    fail(HDLZ_ST_BAD_CODE);                                 // "invalid token" (deflate.py:1559-1560)
                }
            }
            acc >>= used; fill -= used;
            // one append step: the literals, or up to eight bytes of the copy (COPY, deflate.py:1627-1656)
            const bool copying = rem != 0;
            const unsigned long long sv = source8(copying ? dist : 0u);
            const uint32_t m = copying ? (rem < 8u ? rem : 8u) : nlit;
            rem -= copying ? m : 0u;
            if (m && st == HDLZ_OK) append8(copying ? sv : (unsigned long long)lits, m);
        } else if (kDyn && state == S_DYN) {
            // ---- dynamic block: the same trip structure, tables in this lane's global scratch ----
            if (rem == 0) {
                if (fill < 32) {
                    if (wi > nfull + 2) fail(HDLZ_ST_TRUNCATED);
                    else refill();
                }
                if (state == S_DYN) {
                    const uint32_t room = out_cap - o;
                    uint32_t lits = 0, nlit = 0, used = 0;
                    uint32_t e = 0;
                    // up to four literals (or stop at the first other symbol)
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t x = (uint32_t)(acc >> used);
                        e = hot->lit[x & ((1u << kDynLitBits) - 1u)];
                        if ((e & 15u) == 0) e = lane_slow_decode(x, my->cnt_l, my->sorted_l, kDynLitBits, my->resume_l);
                        const uint32_t nb = e & 15u;
                        if (nb == 0 || (e >> 4) >= 256u || used + nb > 32u || nlit >= room) break;
                        lits |= (e >> 4) << (8 * k);
                        used += nb;
                        ++nlit;
                        e = 0xFFFFFFFFu;                                    // consumed
                    }
                    acc >>= used; fill -= used;
                    if (nlit) {
                        append(lits, nlit);
                    } else if ((e & 15u) == 0) {
                        fail(HDLZ_ST_BAD_CODE);                             // no such code
                    } else {
                        const uint32_t nb = e & 15u, info = s_sym[e >> 4];
                        const uint32_t kind = (info >> 8) & 3u;
                        if (kind == 0u) {
                            fail(HDLZ_ST_OUT_OVERFLOW);                     // a literal with no room left
                        } else if (kind == 1u) {
                            acc >>= nb; fill -= nb;
                            state = final_blk ? S_FINISH : S_HEADER;
                            if (final_blk) final_blk = 2;
                        } else if (kind == 3u) {
                            fail(HDLZ_ST_BAD_CODE);
                        } else {
                            const uint32_t eb = (info >> 4) & 15u;
                            if (fill < nb + eb) refill();
                            const uint32_t len = (info >> 16) + ((uint32_t)(acc >> nb) & ((1u << eb) - 1u));
                            acc >>= nb + eb; fill -= nb + eb;
                            if (fill < 32) refill();
                            const uint32_t y = (uint32_t)acc;
                            uint32_t d = hot->dist[y & ((1u << kDynDistBits) - 1u)];
                            if ((d & 15u) == 0) d = lane_slow_decode(y, my->cnt_d, my->sorted_d, kDynDistBits, my->resume_d);
                            const uint32_t dnb = d & 15u;
                            const uint32_t de = s_dsym[(d >> 4) & 31u];
                            const uint32_t deb = de & 15u;
                            if (dnb == 0 || (d >> 4) >= 30u) fail(HDLZ_ST_BAD_CODE);
                            else {
                                const uint32_t dnew = (de >> 8) + ((uint32_t)(acc >> dnb) & ((1u << deb) - 1u));
                                acc >>= dnb + deb; fill -= dnb + deb;        // <= 15 + 13 = 28 bits
                                if (dnew > o) fail(HDLZ_ST_DIST_TOO_FAR);
                                else if (len > room) fail(HDLZ_ST_OUT_OVERFLOW);
                                else { rem = len; dist = dnew; }
                            }
                        }
                    }
                }
            }
            if (state == S_DYN && rem != 0) {
                const uint32_t m = rem < 8u ? rem : 8u;                     // eight bytes per trip, as in fixed blocks
                append8(source8(dist), m);
                rem -= m;
            }
        } else if (state == S_HEADER) {
            if (fill < 32) refill();
            const bool past_end = (uint64_t)wi * 32 - fill + 3 > 8ull * n_in;
            final_blk = (uint32_t)acc & 1u;
            const uint32_t type = past_end ? 4u : ((uint32_t)acc >> 1) & 3u;
            acc >>= 3; fill -= 3;
            if (type == 4) {
                fail(HDLZ_ST_TRUNCATED);                                    // "NO EOF!" (deflate.py:1535-1539)
            } else if (type == 1) {
                state = S_FIXED;
            } else if (type == 3) {
                fail(HDLZ_ST_BAD_BTYPE);                                    // "Bad method" (deflate.py:718-721)
            } else if (type == 0) {
                // stored block header (deflate.py:709-717)
                const uint32_t drop = fill & 7u;
                acc >>= drop; fill -= drop;
                if (fill < 32) refill();
                const uint32_t len = (uint32_t)acc & 0xFFFFu, nlen = ((uint32_t)acc >> 16) & 0xFFFFu;
                acc >>= 32; fill -= 32;
                const uint64_t bytepos = ((uint64_t)wi * 32 - fill) >> 3;
                if ((len ^ 0xFFFFu) != nlen) fail(HDLZ_ST_BAD_STORED);
                else if (bytepos + len > n_in) fail(HDLZ_ST_TRUNCATED);
                else if ((uint64_t)o + len > out_cap) fail(HDLZ_ST_OUT_OVERFLOW);
                else { stored_left = len; state = S_STORED; }
            } else if (!kDyn || !my) {
                // dynamic block: another decoder restarts the stream from its first byte (anything
                // produced so far is simply rewritten) — the dynamic-capable lanes, or without
                // scratch the warp-per-stream kernel
                hand_over = true;
                to_dyn = !kDyn && dyn_list != nullptr;
                state = S_FINISH;
            } else if (kDyn) {
                // ---- dynamic block header (BL / READBL / REPEAT, deflate.py:1084-1202) ----
                auto get = [&](uint32_t n) -> uint32_t {                    // n <= 16
                    if (fill < n) refill();
                    const uint32_t v = (uint32_t)acc & ((1u << n) - 1u);
                    acc >>= n; fill -= n;
                    return v;
                };
                const uint32_t nlen = get(5) + 257, ndist = get(5) + 1, ncode = get(4) + 4;
                uint32_t bad = (nlen > 286 || ndist > 30) ? 1u : 0u;
                uint8_t *lens = my->lens;
                for (int i = 0; i < 19; ++i) lens[i] = 0;
                for (uint32_t i = 0; i < ncode; ++i) lens[c_clorder[i]] = (uint8_t)get(3);
                // code-length code: 7-bit table in the (not yet used) sorted-symbols area of the literal code
                uint16_t *cl_tbl = my->sorted_l;
                if (!bad) bad = lane_build(lens, 19, cl_tbl, 7, my->cnt_d, my->sorted_d, my->resume_d, false);
                if (!bad) {
                    uint32_t any = 0;
                    for (int l = 1; l <= 7; ++l) any |= my->cnt_d[l];
                    if (!any) bad = 1;
                }
                uint8_t *ll = my->lens;                                     // literal/length + distance lengths
                uint32_t idx = 0, prev = 0;
                const uint32_t total = nlen + ndist;
                // the 19 code-length lengths are dead once their table is built: reuse the array
                while (!bad && idx < total) {
                    if (fill < 32) {
                        if (wi > nfull + 2) { bad = 2; break; }
                        refill();
                    }
                    const uint32_t e = cl_tbl[(uint32_t)acc & 127u];
                    const uint32_t nb = e & 15u, sym = e >> 4;
                    if (nb == 0) { bad = 1; break; }
                    acc >>= nb; fill -= nb;
                    uint32_t rep, val;
                    if (sym < 16) { rep = 1; val = sym; prev = sym; }
                    else if (sym == 16) {
                        if (idx == 0) { bad = 1; break; }
                        rep = 3 + get(2); val = prev;
                    } else if (sym == 17) { rep = 3 + get(3); val = 0; prev = 0; }
                    else { rep = 11 + get(7); val = 0; prev = 0; }
                    if (idx + rep > total) { bad = 1; break; }
                    for (uint32_t k = 0; k < rep; ++k) ll[idx + k] = (uint8_t)val;
                    idx += rep;
                }
                if (!bad && ll[256] == 0) bad = 1;                           // no end-of-block code
                if (!bad) bad = lane_build(ll + nlen, (int)ndist, hot->dist, kDynDistBits, my->cnt_d, my->sorted_d, my->resume_d, true);
                if (!bad) bad = lane_build(ll, (int)nlen, hot->lit, kDynLitBits, my->cnt_l, my->sorted_l, my->resume_l, true);
                if (bad) fail(bad == 2 ? HDLZ_ST_TRUNCATED : HDLZ_ST_BAD_CODE);   // "Invalid data" (deflate.py:1140)
                else state = S_DYN;
            }
        } else if (state == S_STORED) {
            // stored bytes, up to 4 per trip (COPY with method 0, deflate.py:1603-1616)
            for (int k = 0; k < 4 && stored_left; ++k, --stored_left) {
                if (fill < 8) refill();
                append((uint32_t)acc & 255u, 1);
                acc >>= 8; fill -= 8;
            }
            if (stored_left == 0) {
                state = final_blk ? S_FINISH : S_HEADER;
                if (final_blk) final_blk = 2;
            }
        } else if (state == S_FINISH) {
            // ---- end of a stream: tail of the output, trailer checks, result words ----
            if (st == HDLZ_OK && !hand_over) {
                if (final_blk != 2) {
                    st = HDLZ_ST_TRUNCATED;
                } else {
                    const uint32_t wo = o >> 2;
                    for (uint32_t x = flushed; x < wo; ++x) dst32[x] = ring[(x & (kRing - 1)) * 32];
                    for (uint32_t k = 0; k < (o & 3u); ++k) dst[(o & ~3u) + k] = (uint8_t)(cw >> (8 * k));
                    const uint64_t bp = (uint64_t)wi * 32 - fill;
                    const uint64_t tp = (bp + 7) >> 3;                       // the trailer (zlib: Adler-32) must be present
                    if (bp > 8ull * n_in || tp + trailer_bytes > n_in) {
                        st = HDLZ_ST_TRUNCATED;                              // "NO EOF!" (deflate.py:1535-1539)
                    } else if (want_crc) {
                        if (crc32_bytes(dst, o, s_nib) != load_le32(src + tp) || o != load_le32(src + tp + 4))
                            st = HDLZ_ST_BAD_CRC;
                    } else if (want_adler) {
                        ad_a %= 65521u; ad_b %= 65521u;
                        const uint32_t want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) |
                                              ((uint32_t)src[tp + 2] << 8) | src[tp + 3];
                        if (((ad_b << 16) | ad_a) != want) st = HDLZ_ST_BAD_ADLER;
                    }
                }
            }
            if (hand_over) {
                if (to_dyn) dyn_list[atomicAdd(dyn_count, 1u)] = (uint32_t)sid;
                else work_list[atomicAdd(work_count, 1u)] = (uint32_t)sid;
            } else {
                out_len[sid] = st == HDLZ_OK ? o : 0;
                if (status) status[sid] = st;
            }
            state = S_IDLE;
        }
    }
}

}  // namespace

int launch_inflate(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                   const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                   uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags, int slot, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    // Few streams: one warp each is the better mapping.  Many streams: one lane each first,
    // the warp-per-stream kernel then finishes whatever was handed over.
    const bool lanes_first = !(flags & HDLZ_F_FORCE_GENERAL) && (n >= 1024 || (flags & HDLZ_F_FORCE_LANES)) &&
                             n < 0xFFFFFFFFull;
    if (!lanes_first)
        return launch_inflate_general(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len,
                                      d_status, n, flags, nullptr, nullptr, s);
    slot %= 3;
    // everything this launch shares with its kernels lives in its slot; the slot's previous launch (any
    // stream) must have finished with it
    if (!ctx->slot_event[slot]) HDLZ_CUDA(cudaEventCreateWithFlags(&ctx->slot_event[slot], cudaEventDisableTiming));
    else HDLZ_CUDA(cudaStreamWaitEvent(s, ctx->slot_event[slot], 0));
    {
        const int rc = grow_device((void **)&ctx->d_workb[slot], &ctx->d_workb_cap[slot], (2 * (size_t)n + 16) * sizeof(uint32_t));
        if (rc) return rc;
    }
    // d_work: [0] count of streams for the warp-per-stream kernel, [1] count of streams with dynamic
    // blocks, [2..3] queue heads of the two-phase route, [16 ..) the two lists (n entries each)
    uint32_t *d_work = ctx->d_workb[slot];
    uint32_t *count = d_work, *dyn_count = d_work + 1, *list = d_work + 16, *dyn_list = d_work + 16 + n;
    HDLZ_CUDA(cudaMemsetAsync(d_work, 0, 4 * sizeof(uint32_t), s));
    uint64_t blocks = (n + kLWarps * 32 - 1) / (kLWarps * 32);
    if (flags & HDLZ_F_PERSISTENT_LANES) {
        const uint64_t resident = (uint64_t)ctx->sm_count * kLaneCtasPerSm;
        if (blocks > resident) blocks = resident;
    }
    // scratch of the lane-per-stream dynamic-block kernel: per-thread tables in global memory
    const uint64_t dyn_blocks = (uint64_t)ctx->sm_count * kDynCtasPerSm;
    const size_t dyn_threads = (size_t)dyn_blocks * (kLWarps * 32);
    const size_t hot_bytes = dyn_threads * sizeof(LaneHot);
    const size_t need = hot_bytes + dyn_threads * sizeof(LaneScratch);
    if (!(flags & HDLZ_F_NO_LANE_SCRATCH) && ctx->d_lane_cap[slot] < need) {
        if (ctx->d_lane[slot]) HDLZ_CUDA(cudaFree(ctx->d_lane[slot]));
        ctx->d_lane[slot] = nullptr;
        ctx->d_lane_cap[slot] = 0;
        if (cudaMalloc(&ctx->d_lane[slot], need) == cudaSuccess) ctx->d_lane_cap[slot] = need;
        else (void)cudaGetLastError();           // no scratch: dynamic streams go to the warp-per-stream kernel
    }
    const bool have = !(flags & HDLZ_F_NO_LANE_SCRATCH) && ctx->d_lane[slot];
    LaneHot *hot = have ? reinterpret_cast<LaneHot *>(ctx->d_lane[slot]) : nullptr;
    LaneScratch *scratch = have ? reinterpret_cast<LaneScratch *>(static_cast<uint8_t *>(ctx->d_lane[slot]) + hot_bytes) : nullptr;
    k_inflate_lanes<false><<<(unsigned)blocks, kLWarps * 32, 0, s>>>(
        d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len, d_status, n, flags, list, count,
        scratch ? dyn_list : nullptr, dyn_count, nullptr, nullptr, 0u, nullptr, nullptr);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    if (scratch) {
        // ---- dynamic-block streams.  Two-phase route (hdlz_inflate_split.cu) for as many of them as its token pool
        // holds; the pool follows the demand the slot's previous launch reported (a batch's dynamic-block count is
        // only known on the device), starting from 256 MiB.
        uint32_t split_items = 0;
        const bool split_ok = !(flags & HDLZ_F_NO_SPLIT) && out_cap <= kSplitMaxOut && out_cap >= 16 &&
                              !((flags & HDLZ_F_GZIP) && (flags & HDLZ_F_VERIFY_ADLER));
        if (split_ok) {
            if (!ctx->h_dyn_seen) {
                if (cudaMallocHost((void **)&ctx->h_dyn_seen, 4 * sizeof(uint32_t)) == cudaSuccess)
                    ctx->h_dyn_seen[0] = ctx->h_dyn_seen[1] = ctx->h_dyn_seen[2] = 0;
                else (void)cudaGetLastError();
            }
            const size_t per = split_slot_bytes(out_cap);
            uint64_t want = (256ull << 20) / per + 1;
            if (ctx->h_dyn_seen && ctx->h_dyn_seen[slot] > want) want = ctx->h_dyn_seen[slot];
            if (want > n) want = n;
            if (ctx->d_split_cap[slot] / per < want) {
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                const size_t budget = (free_b + ctx->d_split_cap[slot]) / 2;          // never more than half of what is free
                if (want * per > budget) want = budget / per;
                if (ctx->d_split_cap[slot] / per < want) {
                    if (ctx->d_split[slot]) HDLZ_CUDA(cudaFree(ctx->d_split[slot]));
                    ctx->d_split[slot] = nullptr;
                    ctx->d_split_cap[slot] = 0;
                    if (cudaMalloc(&ctx->d_split[slot], want * per) == cudaSuccess) ctx->d_split_cap[slot] = want * per;
                    else (void)cudaGetLastError();
                }
            }
            if (!ctx->d_split_scratch[slot]) {
                if (cudaMalloc(&ctx->d_split_scratch[slot], split_scratch_bytes(ctx)) != cudaSuccess) {
                    (void)cudaGetLastError();
                    ctx->d_split_scratch[slot] = nullptr;
                }
            }
            if (ctx->d_split[slot] && ctx->d_split_scratch[slot]) {
                const uint64_t fit = ctx->d_split_cap[slot] / per;
                split_items = (uint32_t)(fit < n ? fit : n);
                const int rc = launch_inflate_split(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap,
                                                    d_out_len, d_status, flags, dyn_list, dyn_count, split_items,
                                                    ctx->d_split[slot], ctx->d_split_scratch[slot], d_work + 2, s);
                if (rc) return rc;
            }
            if (ctx->h_dyn_seen)
                HDLZ_CUDA(cudaMemcpyAsync(ctx->h_dyn_seen + slot, dyn_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        }
        // ---- what the pool did not hold (or the route does not take): the lane-per-stream kernel with per-lane
        // tables in global memory.
        // the primary tables are re-read for every symbol while inputs and outputs stream through the L2
        // once.  With HDLZ_F_PERSIST_TABLES the launch asks the L2 to keep the table range (as much of it
        // as the device lets a window persist).  The carve-out is a device-wide limit and costs the
        // other kernels L2 capacity (the fixed-block kernel lost 10 % with 79 MB set aside), hence
        // opt-in, and given back by the next call without the flag.
        if (!ctx->l2_window) {
            int v = 0;
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            ctx->l2_persist = v;
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            ctx->l2_window = v > 0 ? v : -1;
            (void)cudaGetLastError();
        }
        const bool persist = (flags & HDLZ_F_PERSIST_TABLES) && !split_items && ctx->l2_persist > 0 && ctx->l2_window > 0;
        if (persist != (ctx->l2_carved != 0)) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist ? (size_t)ctx->l2_persist : 0);
            (void)cudaGetLastError();
            ctx->l2_carved = persist ? 1 : 0;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)dyn_blocks);
        cfg.blockDim = dim3(kLWarps * 32);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        unsigned nattr = 0;
        if (persist) {
            const size_t win = hot_bytes < (size_t)ctx->l2_window ? hot_bytes : (size_t)ctx->l2_window;
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = hot;
            attr[0].val.accessPolicyWindow.num_bytes = win;
            attr[0].val.accessPolicyWindow.hitRatio = win <= (size_t)ctx->l2_persist ? 1.0f : (float)ctx->l2_persist / (float)win;
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            nattr = 1;
        }
        cfg.attrs = attr;
        cfg.numAttrs = nattr;
        uint32_t *no_out = nullptr;
        HDLZ_CUDA(cudaLaunchKernelEx(&cfg, k_inflate_lanes<true>, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride,
                                     out_cap, d_out_len, d_status, n, flags, list, count, no_out, no_out,
                                     (const uint32_t *)dyn_list, (const uint32_t *)dyn_count, split_items, scratch, hot));
        ctx->launches++;
        HDLZ_CUDA(cudaGetLastError());
    }
    const int rc = launch_inflate_general(ctx, d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap, d_out_len,
                                          d_status, n, flags, list, count, s);
    if (rc) return rc;
    HDLZ_CUDA(cudaEventRecord(ctx->slot_event[slot], s));
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
