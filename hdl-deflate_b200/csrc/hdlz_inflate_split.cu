// hdlz_inflate_split.cu — two-phase inflater for batches of dynamic-Huffman streams (zlib level 6,
// BASELINE config 4: 100 000 x 32 KiB, OBSIZE = 32768).  Replaces the reference's decode states
// BL / READBL / REPEAT / HF1..HF4 / NEXT / INFLATE / D_NEXT (deflate.py:1084-1517) and COPY
// (:1593-1659) — split where the FSM interleaves them:
//
//   phase 1  k_decode_tokens   one THREAD per stream, decode only.  Symbol decode is serial per stream, so
//            the batch parallelism is across streams; with no window access in this phase a lane never
//            waits for a far back-reference, and its tables are small enough to live in shared memory:
//            an 8-bit literal/length table of 16-bit entries and a 7-bit distance table of 8-bit entries
//            (640 B per lane, interleaved by lane: a bank serves two / four lanes; the counts of the longer
//            codes ride in two 64-bit registers): ten warps per SM.  Longer codes (2 % of the symbols of
//            level-6 streams, but some lane of a warp meets one in every second trip) take a canonical
//            bit-serial search in registers that resumes after the table's bits; the sorted-symbols word it
//            ends in is asked for with cp.async and taken one trip later.  A lane keeps four input words of
//            look-ahead; a trip never moves that window and ends in ONE refill from a four-vector cp.async
//            ring — the same instructions for every lane.  Output: a compact token stream in global
//            memory — literal bytes, and one 32-bit token per copy
//                bits 0..7 literals before the copy | 8..16 copy length (0 = none) | 17..31 distance - 1.
//   phase 2  k_resolve_tokens  one CTA of eight warps per stream, the stream's whole output (<= 32 KiB)
//            staged in shared memory (six streams per SM).  256 tokens per step, one per thread: a block
//            scan gives every token its literal source and its output position; the threads place their
//            literals, then their copies in rounds — a bit per output byte says whether it is final, and a
//            copy runs in the first round in which all it reads is (level-6 copies reach far back: mostly
//            the first; a source that lies before the step's output is final without a look at the bits);
//            a CTA barrier separates a round's readiness checks from its copies and the rounds from each
//            other; copies longer than 32 bytes are done by a whole warp.  The finished window goes
//            to HBM with coalesced 128-bit stores, and the optional Adler-32 is taken from the same reads.
//
// Algorithmic HBM traffic per stream: C bytes read + L bytes written; the token stream adds its
// own write + read (about 0.8 L on level-6 data).
//
// Streams this route does not take (out_cap > 32 KiB, gzip CRC check, pool exhausted) stay on the
// lane-per-stream kernel of hdlz_inflate_lanes.cu.

#include "hdlz_common.cuh"
#include "hdlz_frame.cuh"
#include "hdlz_split.cuh"

namespace hdlz {
namespace {

constexpr int kDecWarps = 5;
constexpr int kDecCtasPerSm = 2;
constexpr int kLitBits = 8;
constexpr int kDistBits = 7;
// per-warp table block: 256 16-bit literal/length entries and 128 8-bit distance entries per lane = 640 B per lane
constexpr int kLitBytes = (1 << kLitBits) * 64;               // entry i of lane l at i * 64 + 2 l
constexpr int kDistBytes = (1 << kDistBits) * 32;             // entry i of lane l at kLitBytes + i * 32 + l (also the code-length-code table)
constexpr int kTabBytes = kLitBytes + kDistBytes;             // 20 KiB per warp
constexpr int kRingSlots = 4;                                 // input ring: four 16-byte vectors per lane
constexpr int kRingBytes = kRingSlots * 512;

constexpr int kResThreads = 256;
constexpr int kResCtasPerSm = 6;
constexpr int kWinBytes = (int)kSplitMaxOut + 16;             // the output of one stream + a zeroed tail to 16 bytes
constexpr int kBitmapBytes = ((int)kSplitMaxOut + 32) / 8 + 12;   // one bit per window byte, a spare word, padded to 16

__constant__ uint8_t c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                      43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_lextra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                     3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                      257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                      12289, 16385, 24577};

// The tables of one lane inside the warp's block.  Literal/length entries are (symbol << 4 | length), kLongCode for a
// code longer than the table (and for an unused index); distance and code-length-code entries are bytes,
// (symbol << 3 | length) with lengths 1..7, 0 for a longer code / unused index.  One multiply-add addresses either;
// a bank serves two (four) lanes.
struct LaneTab {
    uint8_t *lit;          // warp block + 2 * lane
    uint8_t *dist;         // warp block + kLitBytes + lane
    __device__ __forceinline__ uint32_t get_lit(uint32_t idx) const { return *reinterpret_cast<const uint16_t *>(lit + idx * 64u); }
    __device__ __forceinline__ void set_lit(uint32_t idx, uint32_t v) const { *reinterpret_cast<uint16_t *>(lit + idx * 64u) = (uint16_t)v; }
    __device__ __forceinline__ uint32_t get_dist(uint32_t idx) const { return dist[idx * 32u]; }
    __device__ __forceinline__ void set_dist(uint32_t idx, uint32_t v) const { dist[idx * 32u] = (uint8_t)v; }
};

constexpr uint32_t kLongCode = 0xF000u;     // literal/length entry of a code longer than the table: not a literal, length 0

// Canonical Huffman tables of one code (HF1INIT..HF4, SPREAD; deflate.py:1227-1400), built by ONE thread: primary
// table in its shared-memory block (kWide: the 16-bit literal/length table, else the 8-bit one), sorted symbols and
// the resume point of the bit-serial decode in its global scratch, and — packed into 64 bits that stay in a
// register — the number of codes of every length beyond the table (`cbits` bits each), which is all the bit-serial
// decode needs.  Returns 0, 1 for an over-subscribed / illegally incomplete code (zlib's inflate_table rules), 3 for
// a code without symbols where the caller wants one.
template <bool kWide>
__device__ __noinline__ int split_build(const uint8_t *lens, int nsym, LaneTab t, int tbits, int cbits, uint16_t *sorted,
                                        uint16_t *resume, bool allow_incomplete, unsigned long long *beyond)
{
    uint16_t cnt[16], first[16], offs[16], run[16];
    for (int l = 0; l < 16; ++l) { cnt[l] = 0; run[l] = 0; }
    for (int s = 0; s < nsym; ++s) cnt[lens[s]]++;
    cnt[0] = 0;
    unsigned long long pk = 0;
    for (int l = tbits + 1; l <= 15; ++l) pk |= (unsigned long long)cnt[l] << (cbits * (l - tbits - 1));
    *beyond = pk;
    for (int i = 0; i < (1 << tbits); ++i) {
        if (kWide) t.set_lit(i, kLongCode);
        else t.set_dist(i, 0);
    }
    resume[0] = resume[1] = 0;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        const int c = cnt[l];
        left = 2 * left - c;
        if (c) maxlen = l;
        if (left < 0) return 1;
    }
    if (maxlen == 0) return 3;                 // no codes: any use fails later
    if (left > 0 && !(allow_incomplete && maxlen == 1)) return 1;
    uint32_t code = 0, off = 0;
    for (int l = 1; l <= 15; ++l) {
        first[l] = (uint16_t)code;
        offs[l] = (uint16_t)off;
        code = (code + cnt[l]) << 1;
        off += cnt[l];
    }
    resume[0] = tbits < 15 ? first[tbits + 1] : 0;
    resume[1] = tbits < 15 ? offs[tbits + 1] : 0;
    for (int s = 0; s < nsym; ++s) {
        const uint32_t l = lens[s];
        if (!l) continue;
        const uint32_t k = run[l]++;
        sorted[offs[l] + k] = (uint16_t)s;
        if ((int)l <= tbits) {
            const uint32_t rev = __brev(first[l] + k) >> (32 - l);
            for (uint32_t idx = rev; idx < (1u << tbits); idx += 1u << l) {
                if (kWide) t.set_lit(idx, ((uint32_t)s << 4) | l);
                else t.set_dist(idx, ((uint32_t)s << 3) | l);
            }
        }
    }
    return 0;
}

// code longer than the primary table: canonical decode one bit at a time, starting after the `tbits`
// bits the table has already ruled out.  -> (sym << 4) | len, 0 = invalid
__device__ __forceinline__ uint32_t split_slow(uint32_t bits, unsigned long long beyond, int cbits, const uint16_t *sorted,
                                               int tbits, uint32_t resume)
{
    int code = (int)((__brev(bits) >> (32 - tbits)) << 1), first = (int)(resume & 0xFFFFu), index = (int)(resume >> 16);
    for (int l = tbits + 1; l <= 15; ++l) {
        code |= (int)((bits >> (l - 1)) & 1u);
        const int c = (int)((beyond >> (cbits * (l - tbits - 1))) & ((1u << cbits) - 1u));
        if (code - c < first) return ((uint32_t)sorted[index + (code - first)] << 4) | (uint32_t)l;
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0;
}

// The same search without the final look-up: (position in `sorted` << 4) | length, 0 = no such code.  Registers only.
__device__ __forceinline__ uint32_t split_slow_find(uint32_t bits, unsigned long long beyond, int cbits, int tbits, uint32_t resume)
{
    int code = (int)((__brev(bits) >> (32 - tbits)) << 1), first = (int)(resume & 0xFFFFu), index = (int)(resume >> 16);
    for (int l = tbits + 1; l <= 15; ++l) {
        code |= (int)((bits >> (l - 1)) & 1u);
        const int c = (int)((beyond >> (cbits * (l - tbits - 1))) & ((1u << cbits) - 1u));
        if (code - c < first) return ((uint32_t)(index + (code - first)) << 4) | (uint32_t)l;
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0;
}

// literal/length symbol 257..285 -> extra bits (0..5) | base length << 16
__device__ __forceinline__ uint32_t len_entry(uint32_t sym) { return (uint32_t)c_lextra[sym - 257] | ((uint32_t)c_lbase[sym - 257] << 16); }
// distance symbol 0..29 -> extra bits (0..13) | base distance << 8
__device__ __forceinline__ uint32_t dist_entry(uint32_t d) { return (d < 2 ? 0u : (d >> 1) - 1u) | ((uint32_t)c_dbase[d] << 8); }

// ---------------------------------------------------------------------------------------------------
// phase 1: Huffman decode -> token stream
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecWarps * 32, kDecCtasPerSm)
k_decode_tokens(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint64_t in_stride,
                const uint32_t *__restrict__ in_len, uint32_t out_cap, uint32_t flags,
                const uint32_t *__restrict__ items, const uint32_t *__restrict__ item_count, uint32_t max_items,
                SplitScratch *scratch, uint32_t *tokbuf, uint32_t tokcap, uint32_t *litbuf, uint32_t litcap_words,
                uint4 *rec, unsigned int *queue)
{
    const uint32_t n_items = min(*item_count, max_items);
    if (n_items == 0) return;
    extern __shared__ uint32_t s_tab[];                     // [kDecWarps] table blocks, then the input rings
    __shared__ uint32_t s_len[32];                          // length symbol - 257 -> len_entry
    __shared__ uint32_t s_dsym[32];                         // distance symbol -> dist_entry
    __shared__ uint32_t s_slow[kDecWarps * 32];             // per lane: the word of `sorted` a long code asked for (see slow_issue)
    if (threadIdx.x < 29) s_len[threadIdx.x] = len_entry(257 + threadIdx.x);
    if (threadIdx.x < 32) s_dsym[threadIdx.x] = threadIdx.x < 30 ? dist_entry(threadIdx.x) : 0u;
    __syncthreads();

    enum { S_IDLE = 0, S_HEADER = 1, S_BLOCK = 2, S_STORED = 3, S_FINISH = 4, S_DONE = 5 };
    const int lane = threadIdx.x & 31;
    uint8_t *wblock = reinterpret_cast<uint8_t *>(s_tab) + (size_t)(threadIdx.x >> 5) * kTabBytes;
    const LaneTab tab = {wblock + 2 * lane, wblock + kLitBytes + lane};
    unsigned long long beyond_l = 0, beyond_d = 0;          // codes per length beyond the tables (9..15 x 9 bits, 8..15 x 8 bits)
    uint32_t res_l = 0, res_d = 0;                          // where the bit-serial decode resumes: first code | index << 16
    SplitScratch *my = scratch + ((size_t)blockIdx.x * (kDecWarps * 32) + threadIdx.x);
    // input ring of this lane: two slots of four words; slot s at ring[s * 128 .. +4) (16 bytes per lane, 512 per slot)
    uint32_t *ring = s_tab + (size_t)kDecWarps * (kTabBytes / 4) + (size_t)(threadIdx.x >> 5) * (kRingBytes / 4) + 4 * lane;

    const uint32_t trailer_bytes = (flags & HDLZ_F_RAW) ? 0u : (flags & HDLZ_F_GZIP) ? 8u : 4u;
    uint32_t state = S_IDLE;
    uint32_t item = 0, n_in = 0, nfull = 0, tailw = 0;
    const uint8_t *src = in;
    const uint32_t *inw = reinterpret_cast<const uint32_t *>(in);
    uint32_t st = HDLZ_OK;
    uint32_t o = 0, final_blk = 0, stored_left = 0;
    // bit reader: the stream bits from bit `p` of w0 on; w1, w2, w3 follow.  Between trips p < 32; a trip of the
    // block loop reads at most 104 bits past it, so the four words cover every peek of a trip with no refill
    // in between, and ONE refill at the end of the trip — the same instructions for every lane, however many
    // words (0..3) it moves on — replaces the per-symbol `if (p >= 32) advance()` that ran for a few lanes at a
    // time (a quarter of the kernel's issue slots at 5 of 32 lanes, profiles/r02_decode_v4_regions.txt).  The
    // words come out of a four-vector ring in shared memory that cp.async fills three vectors ahead, straight
    // from global memory: the vectors a lane reads were asked for at least one fetch before the newest.
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, wi = 0, p = 0;  // wi = stream word index of w0
    uint32_t rc = 0, vf = 0, v0 = 0, m4 = 0;                 // ring index of w0, vectors fetched (both from vector v0 on), (src & 15) / 4
    uint32_t *tokp = tokbuf, *litp = litbuf;
    uint32_t ntok = 0, litw = 0, litfill = 0, pend = 0;
    uint64_t litacc = 0;
    uint32_t trip = 0;

    auto load_word = [&](int64_t w) -> uint32_t {           // stream word w; zero outside the stream
        uint32_t v = w == (int64_t)nfull ? tailw : 0u;
        if (w >= 0 && w < (int64_t)nfull) v = __ldg(inw + w);
        return v;
    };
    auto fetch_next = [&]() {                                // vector v0 + vf (stream words 4 (v0 + vf) - m4 ..) -> its ring slot
        const int64_t i0 = (int64_t)4 * (v0 + vf) - m4;
        uint32_t *dstw = ring + (vf & (kRingSlots - 1u)) * 128;   // this lane's four words of the slot
        if (i0 >= 0 && i0 + 4 <= (int64_t)nfull) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dstw);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(inw + i0) : "memory");
        } else {
            dstw[0] = load_word(i0); dstw[1] = load_word(i0 + 1); dstw[2] = load_word(i0 + 2); dstw[3] = load_word(i0 + 3);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++vf;
    };
    auto ring_word = [&](uint32_t r) -> uint32_t { return ring[((r >> 2) & (kRingSlots - 1u)) * 128 + (r & 3u)]; };
    // end of a block-loop trip: move on by p / 32 words (0..3), refill the window from the ring, keep the ring
    // three vectors ahead.  No branch except the fetch itself (every fourth word of a lane).
    // (the window is simply read again from the ring at its new place: four loads, no word shuffling)
    auto refill = [&]() {
        // all but the newest fetch have landed: vf - 1 is at least two vectors past the window's first, the window
        // (rc .. rc + 3) reaches into the next vector at most
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const uint32_t n = p >> 5;
        rc += n;
        wi += n;
        p &= 31u;
        w0 = ring_word(rc); w1 = ring_word(rc + 1u); w2 = ring_word(rc + 2u); w3 = ring_word(rc + 3u);
        if (vf < (rc >> 2) + kRingSlots) fetch_next();            // the vector left behind gives its slot to the next one
    };
    // the other states (block headers, stored bytes): one word at a time, when p >= 32
    auto advance = [&]() {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        ++rc;
        ++wi;
        p -= 32u;
        w0 = w1; w1 = w2; w2 = w3;
        w3 = ring_word(rc + 3u);
        if (vf < (rc >> 2) + kRingSlots) fetch_next();
    };
    auto peek = [&]() -> uint32_t { return __funnelshift_r(w0, w1, p); };       // p < 32
    auto peek_at = [&](uint32_t pp) -> uint32_t {                              // pp < 96
        const uint32_t k = pp >> 5;
        return __funnelshift_r(k == 0u ? w0 : k == 1u ? w1 : w2, k == 0u ? w1 : k == 1u ? w2 : w3, pp & 31u);
    };
    // A code longer than its table ends in a look-up in the lane's sorted-symbols array in global memory, and a
    // warp meets one in every second trip: waiting for that load held all 32 lanes for an L2 round trip.  The
    // lane now only ASKS for the word (cp.async into its slot of s_slow) and makes no progress in this trip;
    // the next trip arrives at the same code again and takes the symbol from shared memory, the load having
    // travelled while the other lanes decoded.  slow_kind: 0 nothing asked, 1 literal/length, 2 distance.
    uint32_t slow_kind = 0, slow_meta = 0;                   // meta: code length | (index & 1) << 4
    auto slow_issue = [&](uint32_t found, const uint16_t *sorted) {
        const uint32_t idx = found >> 4;
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&s_slow[threadIdx.x]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\ncp.async.commit_group;"
                     ::"r"(sa), "l"(reinterpret_cast<const uint32_t *>(sorted) + (idx >> 1)) : "memory");
        slow_meta = (found & 15u) | ((idx & 1u) << 4);
    };
    auto slow_take = [&]() -> uint32_t {                     // -> (symbol << 4) | length
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const uint32_t w = *reinterpret_cast<volatile uint32_t *>(&s_slow[threadIdx.x]);
        slow_kind = 0;
        return (((w >> (16u * (slow_meta >> 4))) & 0xFFFFu) << 4) | (slow_meta & 15u);
    };
    auto bitpos = [&]() -> uint64_t { return (uint64_t)wi * 32u + p; };         // stream bits consumed
    auto fail = [&](uint32_t code) { st = code; state = S_FINISH; };
    // append nl (1..3) literal bytes (low bytes of lw)
    auto emit_lits = [&](uint32_t lw, uint32_t nl) {
        litacc |= (uint64_t)lw << (8u * litfill);
        litfill += nl;
        o += nl;
        pend += nl;
        if (litfill >= 4u) {
            litp[litw++] = (uint32_t)litacc;
            litacc >>= 32;
            litfill -= 4u;
        }
        if (pend >= 252u) {                    // a token holds at most 255 literals: close a literal-only one
            tokp[ntok++] = pend;
            pend = 0;
        }
    };

    for (;;) {
        ++trip;
        // ---- idle lanes take their next stream.  Looked at every 16th trip, and only when eight lanes
        // are waiting (or nothing else runs): a lane that opens a stream alone parses the block header
        // and builds its tables with 31 lanes watching.  (Config 4, lanes to wait for: 1: 37.2 ms, 2: 35.4,
        // 4: 34.3, 8: 33.8, 16: 33.6, 32: 34.0 — the wait costs less than the lone table builds.)
        if ((trip & 15u) == 1u) {
            const uint32_t idle = __ballot_sync(HDLZ_FULL_MASK, state == S_IDLE);
            const uint32_t busy = __ballot_sync(HDLZ_FULL_MASK, state != S_IDLE && state != S_DONE);
            if (!idle && !busy) break;
            if (idle && (__popc(idle) >= 8 || !busy)) {
                uint32_t base = 0;
                const int leader = __ffs(idle) - 1;
                if (lane == leader) base = atomicAdd(queue, (unsigned)__popc(idle));
                base = __shfl_sync(HDLZ_FULL_MASK, base, leader);
                if (state == S_IDLE) {
                    item = base + __popc(idle & ((1u << lane) - 1u));
                    if (item >= n_items) {
                        state = S_DONE;
                    } else {
                        const uint64_t sid = items[item];
                        n_in = in_len[sid];
                        src = in + (in_off ? in_off[sid] : sid * in_stride);
                        inw = reinterpret_cast<const uint32_t *>(src);
                        nfull = n_in >> 2;
                        tokp = tokbuf + (size_t)item * tokcap;
                        litp = litbuf + (size_t)item * litcap_words;
                        st = HDLZ_OK;
                        o = 0; ntok = 0; litw = 0; litfill = 0; pend = 0; litacc = 0;
                        slow_kind = 0;
                        final_blk = 0; stored_left = 0;
                        state = S_HEADER;
                        const Frame frame = parse_frame(src, n_in, flags);
                        if (frame.status != HDLZ_OK) {
                            fail(frame.status);
                        } else {
                            tailw = 0;
                            for (uint32_t b = 0; b < (n_in & 3u); ++b) tailw |= (uint32_t)src[4 * nfull + b] << (8 * b);
                            m4 = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u) >> 2;
                            wi = frame.body >> 2;
                            p = 8u * (frame.body & 3u);
                            const uint32_t a = wi + m4;                          // aligned word index of the first word
                            // nothing of the lane's previous stream may still be landing in the ring
                            asm volatile("cp.async.wait_group 0;" ::: "memory");
                            v0 = a >> 2;
                            vf = 0;
                            for (int k = 0; k < kRingSlots; ++k) fetch_next();
                            asm volatile("cp.async.wait_group 0;" ::: "memory");
                            rc = a & 3u;
                            w0 = ring_word(rc); w1 = ring_word(rc + 1u); w2 = ring_word(rc + 2u); w3 = ring_word(rc + 3u);
                        }
                    }
                }
            }
        }

        if (state == S_BLOCK) {
            // ---- NEXT / INFLATE / D_NEXT: up to three literals and then, if one follows, one length/distance
            // pair, per trip.  The common path is branch-free (the lanes of a warp sit in different places of
            // their streams: a branch costs the whole warp both sides); whatever else the next symbol is — end
            // of block, a code longer than the tables, anything invalid — goes through `other` below.
            if (wi > nfull + 4) fail(HDLZ_ST_TRUNCATED);          // far past the end of the input: a runaway decode
            else {
                {
                    const uint32_t x = peek();                              // p < 32 here; nothing below moves the window
                    const uint32_t room = out_cap - o;
                    uint32_t used = 0, nl = 0, lw = 0;
                    bool go = true;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t e = tab.get_lit((x >> used) & ((1u << kLitBits) - 1u));
                        go = go && e < (256u << 4) && nl < room;      // a literal within the table, and room for it
                        lw |= go ? (e >> 4) << (8 * k) : 0u;
                        used += go ? e & 15u : 0u;
                        nl += go ? 1u : 0u;
                    }
                    p += used;
                    // the literals (possibly none) join the literal stream
                    litacc |= (uint64_t)lw << (8u * litfill);
                    litfill += nl;
                    o += nl;
                    pend += nl;
                    if (litfill >= 4u) {
                        litp[litw++] = (uint32_t)litacc;
                        litacc >>= 32;
                        litfill -= 4u;
                    }
                    if (pend >= 252u) {                    // a token holds at most 255 literals: close a literal-only one
                        tokp[ntok++] = pend;
                        pend = 0;
                    }
                }
                const uint32_t x = __funnelshift_r(p < 32u ? w0 : w1, p < 32u ? w1 : w2, p & 31u);   // p < 56
                const uint32_t e = tab.get_lit(x & ((1u << kLitBits) - 1u));
                const uint32_t nb = e & 15u, ls = (e >> 4) - 257u;
                const bool is_len = ls < 29u && nb != 0u;                     // a length symbol of the table
                const uint32_t info = s_len[is_len ? ls : 0u];
                const uint32_t eb = info & 15u;
                const uint32_t len = (info >> 16) + ((x >> nb) & ((1u << eb) - 1u));    // <= 8 + 5 bits of 32
                const uint32_t p2 = p + nb + eb;                                // < 56 + 13
                const uint32_t y = peek_at(p2);
                const uint32_t d = tab.get_dist(y & ((1u << kDistBits) - 1u));
                const uint32_t dnb = d & 7u, dsym = d >> 3;
                const uint32_t de = s_dsym[dsym];
                const uint32_t deb = de & 15u;
                const uint32_t dist = (de >> 8) + ((y >> dnb) & ((1u << deb) - 1u));   // <= 7 + 13 bits of 32
                const bool copy = is_len && dnb != 0u && dsym < 30u && dist <= o && len <= out_cap - o;
                if (copy) {
                    p = p2 + dnb + deb;                                         // < 69 + 20
                    tokp[ntok++] = pend | (len << 8) | ((dist - 1u) << 17);
                    pend = 0;
                    o += len;
                } else if (e >= (256u << 4) || o >= out_cap) {
                    // ---- `other`: not a copy the tables decode, and not a literal the next trip takes
                    uint32_t e2 = e;
                    const bool lit_long = (e2 & 15u) == 0;
                    bool defer = false;                      // asked for a symbol: this trip ends here for the lane
                    if (lit_long) {
                        if (slow_kind == 1u) {
                            e2 = slow_take();
                        } else {
                            const uint32_t found = split_slow_find(x, beyond_l, 9, kLitBits, res_l);
                            if (found) {
                                slow_issue(found, my->sorted_l);
                                slow_kind = 1u;
                                defer = true;
                            }
                        }
                    }
                    const uint32_t nb2 = e2 & 15u, sym = e2 >> 4;
                    if (defer) {
                    } else if (nb2 == 0) {
                        fail(HDLZ_ST_BAD_CODE);                                 // no such code ("Invalid data")
                    } else if (sym < 256u) {
                        if (o >= out_cap) fail(HDLZ_ST_OUT_OVERFLOW);
                        else { p += nb2; emit_lits(sym, 1); }
                    } else if (sym == 256u) {
                        p += nb2;
                        state = final_blk ? S_FINISH : S_HEADER;
                        if (final_blk) final_blk = 2;                           // 2 = finished cleanly
                    } else if (sym > 285u) {
                        fail(HDLZ_ST_BAD_CODE);                                 // "invalid token" (deflate.py:1559-1560)
                    } else {
                        const uint32_t info2 = s_len[sym - 257u];
                        const uint32_t eb2 = info2 & 15u;
                        const uint32_t len2 = (info2 >> 16) + ((x >> nb2) & ((1u << eb2) - 1u));    // <= 15 + 5 bits of 32
                        const uint32_t pd = p + nb2 + eb2;                      // < 56 + 20
                        const uint32_t y2 = peek_at(pd);
                        const uint32_t dq = tab.get_dist(y2 & ((1u << kDistBits) - 1u));
                        uint32_t d2 = ((dq >> 3) << 4) | (dq & 7u);                  // -> (symbol << 4 | length)
                        if ((d2 & 15u) == 0) {
                            if (lit_long) {
                                // both codes longer than their tables: this one waits for its symbol
                                d2 = split_slow(y2, beyond_d, 8, my->sorted_d, kDistBits, res_d);
                            } else if (slow_kind == 2u) {
                                d2 = slow_take();
                            } else {
                                const uint32_t found = split_slow_find(y2, beyond_d, 8, kDistBits, res_d);
                                if (found) {
                                    slow_issue(found, my->sorted_d);
                                    slow_kind = 2u;
                                    defer = true;                               // the length symbol is decoded again next trip
                                }
                            }
                        }
                        const uint32_t dnb2 = d2 & 15u;
                        if (defer) {
                        } else if (dnb2 == 0 || (d2 >> 4) >= 30u) {
                            fail(HDLZ_ST_BAD_CODE);
                        } else {
                            const uint32_t de2 = s_dsym[d2 >> 4];
                            const uint32_t deb2 = de2 & 15u;
                            const uint32_t dist2 = (de2 >> 8) + ((y2 >> dnb2) & ((1u << deb2) - 1u));   // <= 15 + 13 bits
                            p = pd + dnb2 + deb2;
                            if (dist2 > o) fail(HDLZ_ST_DIST_TOO_FAR);           // "distance too big" (deflate.py:1506-1508)
                            else if (len2 > out_cap - o) fail(HDLZ_ST_OUT_OVERFLOW);
                            else {
                                tokp[ntok++] = pend | (len2 << 8) | ((dist2 - 1u) << 17);
                                pend = 0;
                                o += len2;
                            }
                        }
                    }
                }
                refill();                                                   // p < 76 + 28: at most three words
            }
        } else if (state == S_HEADER) {
            const bool past_end = bitpos() + 3 > 8ull * n_in;
            const uint32_t hx = peek();
            final_blk = hx & 1u;
            const uint32_t type = past_end ? 4u : (hx >> 1) & 3u;
            p += 3;
            if (p >= 32u) advance();
            if (type == 4) {
                fail(HDLZ_ST_TRUNCATED);                                        // "NO EOF!" (deflate.py:1535-1539)
            } else if (type == 3) {
                fail(HDLZ_ST_BAD_BTYPE);                                        // "Bad method" (deflate.py:718-721)
            } else if (type == 0) {
                // stored block header (deflate.py:709-717)
                p = (p + 7u) & ~7u;
                if (p >= 32u) advance();
                const uint32_t v = peek();
                const uint32_t len = v & 0xFFFFu, nlen = v >> 16;
                p += 32u;
                advance();
                const uint64_t bytepos = bitpos() >> 3;
                if ((len ^ 0xFFFFu) != nlen) fail(HDLZ_ST_BAD_STORED);
                else if (bytepos + len > n_in) fail(HDLZ_ST_TRUNCATED);
                else if ((uint64_t)o + len > out_cap) fail(HDLZ_ST_OUT_OVERFLOW);
                else { stored_left = len; state = S_STORED; }
            } else {
                uint8_t *lens = my->lens;
                uint32_t bad = 0, nlen = 288, ndist = 32;
                if (type == 1) {
                    // fixed block (STATIC, deflate.py:1064-1076): the same tables, from the fixed lengths (the two
                    // unused 5-bit distance codes are built too and rejected on use, like symbols 286 / 287)
                    for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                    for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
                } else {
                    // ---- dynamic block header (BL / READBL / REPEAT, deflate.py:1084-1202) ----
                    auto get = [&](uint32_t n) -> uint32_t {                    // n <= 16
                        const uint32_t v = peek() & ((1u << n) - 1u);
                        p += n;
                        if (p >= 32u) advance();
                        return v;
                    };
                    nlen = get(5) + 257; ndist = get(5) + 1;
                    const uint32_t ncode = get(4) + 4;
                    bad = (nlen > 286 || ndist > 30) ? 1u : 0u;
                    for (int i = 0; i < 19; ++i) lens[i] = 0;
                    for (uint32_t i = 0; i < ncode; ++i) lens[c_clorder[i]] = (uint8_t)get(3);
                    // code-length code: 7-bit table in the distance table's place, its arrays in the distance code's;
                    // it must be complete and not empty (zlib: "invalid code lengths set")
                    unsigned long long bt = 0;
                    if (!bad) bad = split_build<false>(lens, 19, tab, 7, 8, my->sorted_d, my->resume_d, false, &bt) ? 1u : 0u;
                    uint32_t idx = 0, prev = 0;
                    const uint32_t total = nlen + ndist;
                    while (!bad && idx < total) {
                        if (wi > nfull + 4) { bad = 2; break; }
                        const uint32_t e = tab.get_dist(peek() & 127u);
                        const uint32_t nb = e & 7u, sym = e >> 3;
                        if (nb == 0) { bad = 1; break; }
                        p += nb;
                        if (p >= 32u) advance();
                        uint32_t rep, val;
                        if (sym < 16) { rep = 1; val = sym; prev = sym; }
                        else if (sym == 16) {
                            if (idx == 0) { bad = 1; break; }
                            rep = 3 + get(2); val = prev;
                        } else if (sym == 17) { rep = 3 + get(3); val = 0; prev = 0; }
                        else { rep = 11 + get(7); val = 0; prev = 0; }
                        if (idx + rep > total) { bad = 1; break; }
                        for (uint32_t k = 0; k < rep; ++k) lens[idx + k] = (uint8_t)val;
                        idx += rep;
                    }
                    if (!bad && lens[256] == 0) bad = 1;                         // no end-of-block code
                }
                // an empty distance code is legal (a block of literals only): any use of it fails later
                unsigned long long bd = 0, bl = 0;
                if (!bad) bad = split_build<false>(lens + nlen, (int)ndist, tab, kDistBits, 8, my->sorted_d, my->resume_d, true, &bd) == 1 ? 1u : 0u;
                if (!bad) bad = split_build<true>(lens, (int)nlen, tab, kLitBits, 9, my->sorted_l, my->resume_l, true, &bl) == 1 ? 1u : 0u;
                beyond_d = bd;
                beyond_l = bl;
                res_d = (uint32_t)my->resume_d[0] | ((uint32_t)my->resume_d[1] << 16);
                res_l = (uint32_t)my->resume_l[0] | ((uint32_t)my->resume_l[1] << 16);
                if (bad) fail(bad == 2 ? HDLZ_ST_TRUNCATED : HDLZ_ST_BAD_CODE);   // "Invalid data" (deflate.py:1140)
                else state = S_BLOCK;
            }
        } else if (state == S_STORED) {
            // stored bytes, up to 3 per trip (COPY with method 0, deflate.py:1603-1616); p is a multiple of 8 here
            uint32_t lw = 0, nl = 0;
            const uint32_t v = peek();
            for (int k = 0; k < 3 && stored_left; ++k, --stored_left) {
                lw |= ((v >> (8 * k)) & 255u) << (8 * k);
                ++nl;
            }
            if (nl) {
                p += 8u * nl;
                emit_lits(lw, nl);
                if (p >= 32u) advance();
            }
            if (stored_left == 0) {
                state = final_blk ? S_FINISH : S_HEADER;
                if (final_blk) final_blk = 2;
            }
        } else if (state == S_FINISH) {
            // ---- end of a stream: close the token stream, trailer position, record for phase 2 ----
            uint32_t want = 0;
            if (st == HDLZ_OK) {
                if (final_blk != 2) {
                    st = HDLZ_ST_TRUNCATED;
                } else {
                    if (pend) tokp[ntok++] = pend;
                    if (litfill) litp[litw++] = (uint32_t)litacc;
                    const uint64_t bp = bitpos();
                    const uint64_t tp = (bp + 7) >> 3;                           // the trailer (zlib: Adler-32) must be present
                    if (bp > 8ull * n_in || tp + trailer_bytes > n_in) st = HDLZ_ST_TRUNCATED;   // "NO EOF!" (deflate.py:1535-1539)
                    else if (trailer_bytes == 4u)
                        want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) | ((uint32_t)src[tp + 2] << 8) | src[tp + 3];
                }
            }
            rec[item] = make_uint4(ntok, o, st, want);
            state = S_IDLE;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// phase 2: token stream -> output, one CTA (8 warps) per stream, window in shared memory
// ---------------------------------------------------------------------------------------------------
// bits [start, start + n) of the byte-is-pending bitmap, n <= 32
__device__ __forceinline__ void range_masks(uint32_t start, uint32_t n, uint32_t &w, uint32_t &m0, uint32_t &m1)
{
    w = start >> 5;
    const uint32_t sh = start & 31u;
    const uint32_t full = n >= 32u ? 0xFFFFFFFFu : (1u << n) - 1u;
    m0 = full << sh;
    m1 = sh ? full >> (32u - sh) : 0u;                   // the bits that spill into the next word
}

__device__ __forceinline__ bool range_clear(const uint32_t *bm, uint32_t start, uint32_t n)      // n <= 32
{
    uint32_t w, m0, m1;
    range_masks(start, n, w, m0, m1);
    bool ok = (bm[w] & m0) == 0u;
    if (m1) ok = ok && (bm[w + 1] & m1) == 0u;
    return ok;
}

__device__ __forceinline__ void range_unmark(uint32_t *bm, uint32_t start, uint32_t n)           // n <= 32
{
    uint32_t w, m0, m1;
    range_masks(start, n, w, m0, m1);
    atomicAnd(bm + w, ~m0);
    if (m1) atomicAnd(bm + w + 1, ~m1);
}

__device__ __forceinline__ void range_mark(uint32_t *bm, uint32_t start, uint32_t n)             // n <= 32
{
    uint32_t w, m0, m1;
    range_masks(start, n, w, m0, m1);
    atomicOr(bm + w, m0);
    if (m1) atomicOr(bm + w + 1, m1);
}

__global__ void __launch_bounds__(kResThreads, kResCtasPerSm)
k_resolve_tokens(const uint32_t *__restrict__ items, const uint32_t *__restrict__ item_count, uint32_t max_items,
                 const uint32_t *__restrict__ tokbuf, uint32_t tokcap, const uint32_t *__restrict__ litbuf,
                 uint32_t litcap_words, const uint4 *__restrict__ rec, uint8_t *out, uint64_t out_stride,
                 uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint32_t flags, unsigned int *queue)
{
    const uint32_t n_items = min(*item_count, max_items);
    extern __shared__ uint4 s_win4[];
    uint8_t *win = reinterpret_cast<uint8_t *>(s_win4);                       // the stream's output
    uint32_t *bm = reinterpret_cast<uint32_t *>(win + kWinBytes);             // bit i: byte i of the window waits for a copy of the current step
    __shared__ unsigned long long s_wsum[kResThreads / 32];
    __shared__ unsigned long long s_red[2][kResThreads / 32];
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_adler = (flags & HDLZ_F_VERIFY_ADLER) && !(flags & (HDLZ_F_RAW | HDLZ_F_GZIP));

    for (;;) {
        __syncthreads();                                   // the previous stream's window is no longer read
        if (tid == 0) s_item = atomicAdd(queue, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const uint32_t sid = items[item];
        const uint4 r = rec[item];
        if (r.z != HDLZ_OK) {
            if (tid == 0) {
                out_len[sid] = 0;
                if (status) status[sid] = r.z;
            }
            continue;
        }
        const uint32_t ntok = r.x, o = r.y;
        const uint32_t *tok = tokbuf + (size_t)item * tokcap;
        const uint8_t *lit = reinterpret_cast<const uint8_t *>(litbuf + (size_t)item * litcap_words);
        uint8_t *dst = out + (uint64_t)sid * out_stride;
        for (uint32_t i = tid; i < (kSplitMaxOut + 32) / 32; i += kResThreads) bm[i] = 0;
        __syncthreads();

        uint32_t O0 = 0, L0 = 0;
        for (uint32_t t0 = 0; t0 < ntok; t0 += kResThreads) {
            const uint32_t tk = t0 + tid < ntok ? __ldg(tok + t0 + tid) : 0u;
            const uint32_t lits = tk & 255u, len = (tk >> 8) & 511u, dist = (tk >> 17) + 1u;
            // block scan of both cursors: literals consumed in the low half, bytes produced in the high half
            const unsigned long long packed = (unsigned long long)lits | ((unsigned long long)(lits + len) << 32);
            unsigned long long incl = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long v = __shfl_up_sync(HDLZ_FULL_MASK, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            unsigned long long before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kResThreads / 32; ++w) {
                const unsigned long long v = s_wsum[w];
                if (w < warp) before += v;
                total += v;
            }
            const unsigned long long excl = before + incl - packed;
            const uint32_t lx = (uint32_t)excl;              // literals before this token (within the step)
            const uint32_t ooff = O0 + (uint32_t)(excl >> 32);
            const uint32_t cd = ooff + lits;                 // destination of this thread's copy
            const uint32_t cs = cd - dist;
            const uint32_t need = len < dist ? len : dist;   // an overlapping copy only reads `dist` bytes it did not write
            // every byte a copy of this step will write is marked pending; what the bitmap does not mark is final
            // (literals included, once the barrier below has passed)
            for (uint32_t b = 0; b < len; b += 32) range_mark(bm, cd + b, min(32u, len - b));
            // ---- literals of the warp's 32 tokens, one byte per lane per pass: they are contiguous in the
            // literal stream, the destination comes from the token that owns the byte (binary search over the
            // warp's prefix)
            {
                const uint32_t l0 = __shfl_sync(HDLZ_FULL_MASK, lx, 0);
                const uint32_t l1 = __shfl_sync(HDLZ_FULL_MASK, lx + lits, 31);
                const uint8_t *lsw = lit + L0 + l0;
                for (uint32_t g0 = 0; g0 < l1 - l0; g0 += 32) {          // warp-uniform trip count: the shuffles need all lanes
                    const uint32_t g = g0 + lane;
                    const bool act = g < l1 - l0;
                    const uint32_t v = act ? lsw[g] : 0u;
                    // last token t with lx[t] - l0 <= g
                    uint32_t t = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t probe = __shfl_sync(HDLZ_FULL_MASK, lx, (t + step) & 31);
                        if (probe - l0 <= g) t += step;
                    }
                    const uint32_t tl = __shfl_sync(HDLZ_FULL_MASK, lx, t);
                    const uint32_t to = __shfl_sync(HDLZ_FULL_MASK, ooff, t);
                    if (act) win[to + (g - (tl - l0))] = (uint8_t)v;
                }
            }
            __syncthreads();
            // ---- copies (COPY, deflate.py:1627-1656): a copy runs in the first round in which nothing it reads
            // is pending.  Rounds are separated by a CTA barrier; their number is the depth of the longest
            // chain of copies reading each other inside these 256 tokens (a handful).
            bool pending = len != 0u;
            for (;;) {
                bool ready = false;
                if (pending) {
                    ready = true;
                    // what lies before this step's output is final; only a source that reaches into the step is checked
                    if (cs + need > O0)
                        for (uint32_t b = 0; b < need && ready; b += 32) ready = range_clear(bm, cs + b, min(32u, need - b));
                }
                // No copy of a round starts before every readiness check of the round is done: a thread that checked
                // late could otherwise see bits a copy of the SAME round had already cleared and read that copy's bytes
                // in the same round — ordered by the fence below, but a hand-over compute-sanitizer's racecheck (which
                // knows only barriers) reports as a hazard on the window.  The barrier costs 1 % (33.8 -> 34.2 ms on
                // config 4) and makes the route racecheck-clean (profiles/r02_sanitizer.txt).
                __syncthreads();
                if (__any_sync(HDLZ_FULL_MASK, pending)) {          // a warp with nothing left only keeps the barrier
                // short copies that do not overlap their source: all their bytes as one list, a byte per lane per pass
                {
                    const bool flat = ready && len <= 32u && dist >= len;
                    const uint32_t fl = flat ? len : 0u;
                    uint32_t fi = fl;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(HDLZ_FULL_MASK, fi, d);
                        if (lane >= d) fi += v;
                    }
                    const uint32_t fx = fi - fl;             // bytes of the list before this token
                    const uint32_t ftot = __shfl_sync(HDLZ_FULL_MASK, fi, 31);
                    for (uint32_t g0 = 0; g0 < ftot; g0 += 32) {            // warp-uniform trip count
                        const uint32_t g = g0 + lane;
                        uint32_t t = 0;
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
                            const uint32_t probe = __shfl_sync(HDLZ_FULL_MASK, fx, (t + step) & 31);
                            if (probe <= g) t += step;
                        }
                        // tokens without bytes share their successor's prefix: the search lands on the last of
                        // them, the owner is the first token at or after it that has bytes... the prefix is
                        // non-decreasing, so take the LAST token whose prefix is <= g: it owns g iff it has bytes,
                        // and it does, because a token without bytes has the same prefix as its successor
                        const uint32_t tx = __shfl_sync(HDLZ_FULL_MASK, fx, t);
                        const uint32_t td = __shfl_sync(HDLZ_FULL_MASK, cd, t);
                        const uint32_t ts = __shfl_sync(HDLZ_FULL_MASK, cs, t);
                        if (g < ftot) win[td + (g - tx)] = win[ts + (g - tx)];
                    }
                }
                // short copies that overlap their source: byte-serial by their thread
                if (ready && len <= 32u && dist < len)
                    for (uint32_t b = 0; b < len; ++b) win[cd + b] = win[cs + b];
                // long copies: the whole warp, 32 bytes per step
                uint32_t longs = __ballot_sync(HDLZ_FULL_MASK, ready && len > 32u);
                while (longs) {
                    const int src_lane = __ffs(longs) - 1;
                    longs &= longs - 1;
                    const uint32_t fd = __shfl_sync(HDLZ_FULL_MASK, cd, src_lane);
                    const uint32_t fdist = __shfl_sync(HDLZ_FULL_MASK, dist, src_lane);
                    const uint32_t flen = __shfl_sync(HDLZ_FULL_MASK, len, src_lane);
                    if (fdist >= 32u) {
                        for (uint32_t k = 0; k < flen; k += 32) {
                            if (k + lane < flen) win[fd + k + lane] = win[fd - fdist + k + lane];
                            __syncwarp();
                        }
                    } else {
                        // the source is one period of `fdist` final bytes
                        for (uint32_t k = lane; k < flen; k += 32) win[fd + k] = win[fd - fdist + k % fdist];
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (ready) {
                    __threadfence_block();                   // a thread still checking this round may already see the bits
                    for (uint32_t b = 0; b < len; b += 32) range_unmark(bm, cd + b, min(32u, len - b));
                    pending = false;
                }
                }
                if (!__syncthreads_or(pending)) break;
            }
            O0 += (uint32_t)(total >> 32);
            L0 += (uint32_t)total;
        }

        // ---- the finished window -> HBM (coalesced 128-bit stores), Adler-32 from the same reads
        for (uint32_t k = o + tid; k < ((o + 15u) & ~15u); k += kResThreads) win[k] = 0;
        __syncthreads();
        const uint4 *w4 = reinterpret_cast<const uint4 *>(win);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        const uint32_t nvec = (o + 15u) >> 4;
        unsigned long long s1 = 0, s2 = 0;
        for (uint32_t v = tid; v < nvec; v += kResThreads) {
            const uint4 q = w4[v];
            if (16u * v + 16u <= o) {
                d4[v] = q;
            } else {
                for (uint32_t b = 16u * v; b < o; ++b) dst[b] = win[b];
            }
            if (want_adler) {
                // sum x and sum (o - i) x over the 16 bytes at i = 16 v ..: (o - 16 v) * S - sum j x_j
                const uint32_t S = __dp4a(q.x, 0x01010101u, __dp4a(q.y, 0x01010101u, __dp4a(q.z, 0x01010101u, __dp4a(q.w, 0x01010101u, 0u))));
                const uint32_t J = __dp4a(q.x, 0x03020100u, __dp4a(q.y, 0x07060504u, __dp4a(q.z, 0x0B0A0908u, __dp4a(q.w, 0x0F0E0D0Cu, 0u))));
                s1 += S;
                s2 += (unsigned long long)(o - 16u * v) * S - J;
            }
        }
        uint32_t stt = HDLZ_OK;
        if (want_adler) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                s1 += __shfl_xor_sync(HDLZ_FULL_MASK, s1, d);
                s2 += __shfl_xor_sync(HDLZ_FULL_MASK, s2, d);
            }
            if (lane == 0) { s_red[0][warp] = s1; s_red[1][warp] = s2; }
            __syncthreads();
            if (tid == 0) {
                unsigned long long t1 = 0, t2 = 0;
                for (int w = 0; w < kResThreads / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
                const uint32_t a = (uint32_t)((1ull + t1) % 65521ull);
                const uint32_t b = (uint32_t)(((unsigned long long)o + t2) % 65521ull);
                if (((b << 16) | a) != r.w) stt = HDLZ_ST_BAD_ADLER;
            }
        }
        if (tid == 0) {
            out_len[sid] = stt == HDLZ_OK ? o : 0;
            if (status) status[sid] = stt;
        }
    }
}

}  // namespace

size_t split_slot_bytes(uint32_t out_cap)
{
    return (size_t)split_tokcap(out_cap) * 4 + (size_t)split_litcap_words(out_cap) * 4 + sizeof(uint4);
}

size_t split_scratch_bytes(const hdlz_ctx *ctx)
{
    return (size_t)ctx->sm_count * kDecCtasPerSm * (kDecWarps * 32) * sizeof(SplitScratch);
}

// Runs both phases over the first min(*d_item_count, max_items) entries of d_items.  `pool` holds max_items
// slots (split_slot_bytes each: records | tokens | literals), `scratch` split_scratch_bytes, `queues` two
// zeroed counters.
int launch_inflate_split(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                         const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                         uint32_t *d_out_len, uint32_t *d_status, uint32_t flags, const uint32_t *d_items,
                         const uint32_t *d_item_count, uint32_t max_items, void *pool, void *scratch,
                         unsigned int *queues, cudaStream_t s)
{
    if (max_items == 0) return HDLZ_SUCCESS;
    const uint32_t tokcap = split_tokcap(out_cap), litw = split_litcap_words(out_cap);
    uint4 *rec = reinterpret_cast<uint4 *>(pool);
    uint32_t *tokbuf = reinterpret_cast<uint32_t *>(rec + max_items);
    uint32_t *litbuf = tokbuf + (size_t)max_items * tokcap;
    const size_t dec_smem = (size_t)kDecWarps * (kTabBytes + kRingBytes);
    const size_t res_smem = (size_t)kWinBytes + kBitmapBytes;
    if (!ctx->split_attr_set) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_decode_tokens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
        HDLZ_CUDA(cudaFuncSetAttribute(k_decode_tokens, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_resolve_tokens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
        HDLZ_CUDA(cudaFuncSetAttribute(k_resolve_tokens, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->split_attr_set = true;
    }
    k_decode_tokens<<<(unsigned)(ctx->sm_count * kDecCtasPerSm), kDecWarps * 32, dec_smem, s>>>(
        d_in, d_in_off, in_stride, d_in_len, out_cap, flags, d_items, d_item_count, max_items,
        reinterpret_cast<SplitScratch *>(scratch), tokbuf, tokcap, litbuf, litw, rec, queues);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    k_resolve_tokens<<<(unsigned)(ctx->sm_count * kResCtasPerSm), kResThreads, res_smem, s>>>(
        d_items, d_item_count, max_items, tokbuf, tokcap, litbuf, litw, rec, d_out, out_stride, d_out_len, d_status,
        flags, queues + 1);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
