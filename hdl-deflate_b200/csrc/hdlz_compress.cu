// hdlz_compress.cu — deflate compressor, in its fixed-tree mode bit-exact with the reference's
// FAST + MATCH10 / CWINDOW=32 engine (deflate.py states CSTATIC, SEARCH, SEARCHF,
// DISTANCE, CHECKSUM; :734-1016), re-designed for sm_100a.  Not a port of the FSM:
//
//   persistent warps, one WARP per stream at a time, tiles of 1024 input positions;
//   everything between the HBM read of the input and the HBM write of the stream lives
//   in shared memory and registers.  Streams are handed out through a device-wide queue
//   head (one atomic per stream), not a static stride.
//
//   load     HBM -> shared with 128-bit cp.async (LDGSTS), 32 bytes of history and 32 bytes
//            of look-ahead around the tile.
//   phase A  lane <-> position (32 positions per step)
//            R[p] = 32-bit mask, bit (32-d) <=> x[p-d] == x[p], d = 1..32.
//            A 256-entry per-warp table T[value] = lane mask of that value: the lanes read
//            the previous chunk's entry, clear the previous chunk's entries, OR their own bit
//            in with a shared-memory atomic and read the entry back; one funnel shift joins
//            both halves.  (Not MATCH.ANY: it costs 2 cycles of the SM-wide ADU pipe per
//            distinct lane value and bound the whole kernel, profiles/r01_ubench_match.txt.)
//            This replaces the 32 `matcher3` comparators + cwindow shift register
//            (deflate.py:407-421, 442-453).  R[q] = 0 for q >= L-2 encodes all the
//            `di < isize - k` guards of SEARCH/SEARCHF (deflate.py:913-952, 975-977).
//            Adler-32 partial sums ride along (CSTATIC/CHECKSUM, :826-831, :884-897).
//   phase B  lane <-> segment of 32 consecutive positions, walked from the back
//            M3 = R[p] & R[p+1] & R[p+2]  -> 3-byte match at every distance at once;
//            nearest distance = highest set bit (lowest `si` first, deflate.py:982-988);
//            length = 3 + leading run of that bit through R[p+3..p+9] (SEARCHF), counted
//            by summing the surviving bit in a 64-bit accumulator;
//            token bits from two small LUTs (fixed Huffman, DISTANCE :836-882).
//            The same backward walk runs the parse DP: h[j] = skip count left for
//            the next segment if a token starts at j (10-nibble shift register).
//   phase P2 resolve the entry skip count of each of the 32 segments (greedy parse
//            `di += match` / `di += 1`, deflate.py:960, 1008) with 32 broadcast loads of the published maps.
//   phase P3 ONE forward walk per lane: the tokens that start in the segment are
//            concatenated into a lane-private bitstream (put/do_flush, deflate.py:535-567);
//            a warp scan of the bit counts gives every lane its offset, and a merge pass
//            shifts the private words into the tile's stream (interior words plain stores,
//            the two boundary words OR-ed).
//   flush    whole words of the staged stream go to HBM with coalesced stores; the
//            last tile appends EOB, pad and Adler-32 (deflate.py:771-814).
//
// Algorithmic HBM traffic per stream: L bytes read + C bytes written (C = stream size).
//
// Two more instantiations share everything up to the parse (so the tokens stay the reference's) and differ in
// what P3 does with a token (hdlz_tree.cu): kModeTree codes it with the tree of the context
// instead of the fixed one — one BTYPE = 10 block per stream, opened by a description of the code that is the
// same bit string for every stream; kModeHist only counts the symbols (hdlz_train_tree).

#include "hdlz_common.cuh"

#include <type_traits>

namespace hdlz {
namespace {

constexpr int kTile = 1024;            // positions per tile
constexpr int kSeg = 32;               // positions per lane in the segment phases
constexpr int kChunks = kTile / 32;
constexpr int kWarpsPerCta = 4;
constexpr int kCtasPerSm = 8;
constexpr int kInBytes = 32 + kTile + 32;              // history | tile | look-ahead chunk
// a lane's private stream holds the tokens that START in its segment: at most 31 nine-bit literals
// plus one 15-bit match token at the last position = 294 bits
// With an application / trained code (kMode 1, hdlz_tree.cu) a literal costs up to 15 bits and a match up to
// 15 + 15 + 3: the private streams and the tile's stream are sized for that.
constexpr int kModeFixed = 0, kModeTree = 1, kModeHist = 2;
__host__ __device__ constexpr int priv_words(int mode) { return mode == kModeTree ? ((kSeg - 1) * 15 + 33 + 31) / 32 : ((kSeg - 1) * 9 + 15 + 31) / 32; }   // 16 : 10
__host__ __device__ constexpr int out_words(int mode) { return mode == kModeTree ? 488 : 296; }   // 31 carry bits + 1024 * (15 : 9) + EOB + trailer, rounded up
constexpr int kRWords = (kTile + 32) + (kTile + 32) / kSeg;   // padded: idx = i + i/32  (1089 incl. last +1)

template <int kMode>
struct __align__(16) WarpSmemT {
    static constexpr int kStageBytes = priv_words(kMode) * 32 * 4;   // 1280 (2048) >= kInBytes: input tile, later the private streams
    uint8_t stage[kStageBytes];                        // input tile, then the segment maps, then the lane-private streams
    uint32_t R[(kRWords + 4) / 4 * 4];                 // masks, overwritten in place by tokens; after P3 the
                                                       // tile's part of the output stream (out_words words)
    uint32_t T[256];                                   // value -> lane mask of the previous chunk
    static_assert(kStageBytes >= kInBytes, "stage buffer must hold the input tile");
    static_assert(out_words(kMode) <= kRWords, "the output stream part must fit in the token array");
};

constexpr int kLutWords = kTreeLutWords;   // LT[256] literal tokens, DC[32 + 4] distance part of match tokens (by mask bit), 8 length symbols (tree)
static_assert(kLutWords >= 256 + 36, "fixed-code tables");
template <int kMode>
constexpr size_t smem_bytes() { return kLutWords * 4 + sizeof(WarpSmemT<kMode>) * kWarpsPerCta; }

// token word: bits 0..14 code (LSB-first), 16..19 bit count, 24..29 = 4 * (length in positions - 1),
// i.e. the nibble shift of the parse DP
__device__ __forceinline__ uint32_t rev_n(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// distance part of a match token for distance d = cl + 1: (5-bit code + extra bits) << 7,
// bit count 12 + extra (7 for the length symbol), length field preset to 3
__device__ __forceinline__ uint32_t dist_token_entry(int cl)
{
    // CopyDistance / ExtraDistanceBits (deflate.py:106-110): code c, extra bits eb
    uint32_t e = (uint32_t)cl, c, eb, extra;
    if (e < 4) {
        c = e; eb = 0; extra = 0;
    } else {
        uint32_t msb = 31 - __clz(e);
        eb = msb - 1;
        c = 2 * msb + ((e >> eb) & 1);
        extra = e & ((1u << eb) - 1);
    }
    const uint32_t dc = rev_n(c, 5) | (extra << 5);
    return (dc << 7) | ((12 + eb) << 16) | (8u << 24);
}

__device__ __forceinline__ uint32_t literal_token_entry(uint32_t x)
{
    // fixed Huffman literal codes, bit-reversed (== out_codes[x], deflate.py:112-149); length 1
    return (x < 144 ? (rev_n(0x30 + x, 8) | (8u << 16)) : (rev_n(0x100 + x, 9) | (9u << 16)));
}

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n)      // n >= 32 -> 0 (PTX shl semantics)
{
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}

__device__ __forceinline__ unsigned long long shr64_clamp(unsigned long long v, uint32_t n)   // n >= 64 -> 0
{
    unsigned long long r;
    asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(n));
    return r;
}

__device__ __forceinline__ uint32_t find_msb(uint32_t v)                    // index of the highest set bit, 0xFFFFFFFF for 0
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ unsigned long long add_wide(uint32_t a, unsigned long long c)   // c + a on the FMA pipe
{
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(d) : "r"(a), "l"(c));
    return d;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// kMaxMatch: 10 = the reference's MATCH10 configuration, 5 = MATCH10 False (deflate.py:34-35, 913-924)
// kStream: one stream fed in pieces (hdlz_stream_*): the launch works on the tiles [ctl->t0, ctl->t_end) of a
// stream whose first `uniform_len` bytes have arrived, resumes from and saves to *ctl what the reference's FSM
// keeps between clocks — bit cursor and partial output word (`do` / `doo` / `ob1`, deflate.py:535-567), the parse
// position (`di`), both Adler sums — and writes the words it completes to out[0 ..]; `in` is the address the
// stream's byte 0 would have (only bytes from t0 - 32 on are read).  The last launch (ctl->final) knows the true
// length and closes the stream.  Compiled out of the batch kernel.
// kMode: kModeFixed = the reference's fixed code; kModeTree = the code of *tree (hdlz_set_tree / hdlz_train_tree): same
// parse, every stream starts with tree->prefix, tokens come from tree->lut; kModeHist = no output, the symbols the
// parse produces are counted into hist[] (hdlz_train_tree).
// kLong: ONE long stream (`uniform_len` bytes at `in`) spread over the whole grid, a tile of 1024 positions per
// warp at a time, tiles handed out in order by the queue head.  What the reference's FSM carries from position to
// position crosses tile borders through `lbuf` (three arrays of one entry per tile, zeroed by the launcher):
//   - the parse (`di += match`, deflate.py:960): a tile publishes its MAP carry-in -> carry-out (ten nibbles; the
//     greedy parse forgets where it started within a few tokens, so the map is almost always constant and the
//     carry-out known at once) and finds its own carry-in by looking back over the maps of its predecessors;
//   - the bit cursor (`do` / `doo`, deflate.py:535-567): a decoupled look-back over the tiles' bit counts gives a
//     tile its place in the stream; the tile's words go straight there, the two words it shares with its
//     neighbours by atomic OR into the zeroed output;
//   - the Adler-32 sums (deflate.py:826-831): per tile, folded together by the last tile (adler32_combine rule).
// The bytes are those of the one-warp kernel (tests/test_gpu_compress.py::test_long_stream_over_the_grid).
constexpr unsigned long long kLbResolved = 2ull << 62, kLbPartial = 1ull << 62, kLbValue = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long lb_poll(const unsigned long long *p)     // spin until the entry is published
{
    unsigned long long v;
    do {
        v = *reinterpret_cast<const volatile unsigned long long *>(p);
    } while (v == 0ull);
    return v;
}

template <int kMaxMatch, bool kStream, int kMode, bool kLong = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm)
k_compress(const uint8_t *__restrict__ in, uint64_t in_stride, const uint32_t *__restrict__ in_len,
           uint32_t uniform_len, uint8_t *__restrict__ out, uint64_t out_stride,
           uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams, unsigned long long *queue,
           uint32_t container, StreamCtl *ctl, const TreeDev *__restrict__ tree, unsigned long long *hist,
           unsigned long long *lbuf = nullptr)
{
    static_assert(!kLong || (!kStream && kMode != kModeHist), "the long-stream mode: batch calls, fixed or installed tree");
    extern __shared__ uint4 smem_raw[];
    uint32_t *LT = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *DC = LT + 256;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    using WarpSmem = WarpSmemT<kMode>;
    constexpr int kOutWords = out_words(kMode);
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(LT + kLutWords)[warp];

    if constexpr (kMode == kModeFixed) {
        for (int i = threadIdx.x; i < 256; i += kWarpsPerCta * 32) LT[i] = literal_token_entry((uint32_t)i);
        if (threadIdx.x < 36) DC[threadIdx.x] = threadIdx.x < 32 ? dist_token_entry(31 - (int)threadIdx.x) : 0u;   // by mask bit: d = 32 - f
    } else if constexpr (kMode == kModeTree) {
        for (int i = threadIdx.x; i < kTreeLutWords; i += kWarpsPerCta * 32) LT[i] = tree->lut[i];
    } else {
        for (int i = threadIdx.x; i < kTreeHistWords; i += kWarpsPerCta * 32) LT[i] = 0;      // the CTA's counts
    }
    __syncthreads();

    uint8_t *in_s = ws.stage;                                  // byte i <-> position t0 - 32 + i
    uint32_t *priv = reinterpret_cast<uint32_t *>(ws.stage);   // word k of lane l at priv[k * 32 + l]
    uint32_t *Rw = ws.R;
    uint32_t *T = ws.T;
    uint32_t *outw = ws.R;                                      // valid only after P3 (tokens are dead then)

    for (int i = lane; i < 256; i += 32) T[i] = 0;
    uint32_t vprev = 0;
    __syncwarp();

    // Work distribution: every warp starts on stream <its global index>, later streams come from a
    // device-wide queue head (one atomic per stream).  A static stride would leave the SM idle
    // wherever a CTA of the grid was not resident from the start and ran after the others.
    const uint64_t n_warps = kLong ? 0ull : (uint64_t)gridDim.x * kWarpsPerCta;
    // kLong: `sid` numbers the TILES of the streams (stream-major: all streams have `uniform_len` bytes) and every
    // tile comes from the queue head, the first one too: a tile may wait for its predecessors in its stream, so
    // they must belong to warps that already run
    // kLong with `ctl`: one PIECE of a stream fed in pieces (hdlz_cstream_*): the tiles [ctl->t0, ctl->t_end) — or
    // to the end of the stream when ctl->final — of which `uniform_len` bytes have arrived; tile 0 takes carry,
    // partial word and Adler sums from *ctl, the last tile leaves them there (the partial word stays in `out`).
    uint32_t pc_t0 = 0, pc_final = 1;
    uint64_t n_tiles = kLong ? ((uint64_t)uniform_len + kTile - 1) / kTile : 1ull;     // per stream
    if (kLong && ctl) {
        pc_t0 = ctl->t0;
        pc_final = ctl->final;
        n_tiles = pc_final ? ((uint64_t)(uniform_len - pc_t0) + kTile - 1) / kTile : (ctl->t_end - pc_t0) / kTile;
    }
    const uint64_t all_tiles = n_tiles * n_streams;
    uint64_t sid0 = (uint64_t)blockIdx.x * kWarpsPerCta + warp;
    if constexpr (kLong) {
        unsigned long long tk0 = 0;
        if (lane == 0) tk0 = atomicAdd(queue, 1ull);
        sid0 = __shfl_sync(HDLZ_FULL_MASK, tk0, 0);
    }
    for (uint64_t sid = sid0; sid < (kLong ? all_tiles : n_streams);) {
        unsigned long long next_ticket = 0;
        if (lane == 0) next_ticket = atomicAdd(queue, 1ull);           // in flight while this stream is processed
        const uint64_t ls = kLong ? sid / n_tiles : sid;               // the stream, and (kLong) the tile in it
        const uint64_t lt = kLong ? sid - ls * n_tiles : 0ull;
        // the stream's look-back arrays: one entry per tile
        unsigned long long *lb_map = lbuf + ls * n_tiles, *lb_bits = lbuf + all_tiles + ls * n_tiles;
        uint32_t *lb_adler = reinterpret_cast<uint32_t *>(lbuf + 2 * all_tiles) + ls * n_tiles;
        uint32_t *lb_flag = reinterpret_cast<uint32_t *>(lbuf + 2 * all_tiles) + all_tiles + ls;      // kModeTree: a tile met a symbol without a code
        const uint32_t L = kLong ? uniform_len : in_len ? in_len[sid] : uniform_len;
        const uint8_t *src = in + ls * in_stride;
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(out + ls * out_stride);
        uint64_t need = 0;               // slot size this stream may need
        if constexpr (kMode == kModeFixed) need = compress_bound(L, container);
        else if constexpr (kMode == kModeTree)
            need = ((((uint64_t)tree->prefix_bits + (uint64_t)L * tree->worst_bits + (tree->eob >> 16) + 7) >> 3) +
                    (container == HDLZ_CONTAINER_GZIP ? 8u : container == HDLZ_CONTAINER_RAW ? 0u : 4u) + 15) & ~15ull;
        if (!kStream && !kLong && (L < HDLZ_MIN_INPUT || need > out_stride)) {
            if (lane == 0 && kMode != kModeHist) {
                out_len[sid] = 0;
                if (status) status[sid] = L < HDLZ_MIN_INPUT ? HDLZ_ST_SHORT_INPUT : HDLZ_ST_OUT_OVERFLOW;
            }
            sid = n_warps + __shfl_sync(HDLZ_FULL_MASK, next_ticket, 0);
            continue;
        }

        uint32_t carry = 0;              // positions of the next tile still covered by the last token
        uint32_t adler_a = 1, adler_b = 0;
        // partial output word: container header + BFINAL/BTYPE=01 (zlib: 78 9C, deflate.py:753-761)
        uint32_t pw = 0x78u | (0x9Cu << 8) | (3u << 16);
        uint32_t lbit = 19;              // valid bits in pw
        uint32_t wbase = 0;              // 32-bit words of the stream already written to HBM
        uint32_t uncoded = 0;            // kModeTree: a token whose symbol has no code in the tree
        if constexpr (kMode == kModeTree) {
            // container header, BFINAL = 1 / BTYPE = 10 and the code description: the same bits in every stream
            const uint32_t nfp = tree->prefix_bits >> 5;
            if (!kLong || lt == 0)
                for (uint32_t k = lane; k < nfp; k += 32) dst32[k] = tree->prefix[k];
            pw = tree->prefix[nfp];
            lbit = tree->prefix_bits & 31u;
            wbase = nfp;
        } else if constexpr (kMode == kModeHist) {
            // nothing is written
        } else if (container == HDLZ_CONTAINER_RAW) {
            pw = 3u;
            lbit = 3;
        } else if (container == HDLZ_CONTAINER_GZIP) {
            // 1F 8B 08 00 | MTIME 0 | XFL 0, OS FF (RFC 1952): eight bytes go out now, two ride in pw
            if (lane == 0) {
                dst32[0] = 0x00088B1Fu;
                dst32[1] = 0u;
            }
            pw = (0xFFu << 8) | (3u << 16);
            wbase = 2;
        }

        uint32_t t_first = 0, t_stop = L;
        bool closing = true;             // this launch reaches the end of the stream
        if constexpr (kLong) {           // exactly the tile `lt`
            t_first = pc_t0 + (uint32_t)lt * kTile;
            t_stop = t_first + kTile < L ? t_first + kTile : L;
            closing = pc_final != 0;
            if (lt == 0 && pc_t0 != 0) {                 // a later piece: the partial word the previous one left
                pw = ctl->pw;
                lbit = ctl->lbit;
                wbase = 0;
            }
        }
        if (kStream) {
            t_first = ctl->t0;
            closing = ctl->final != 0;
            if (!closing) t_stop = ctl->t_end;
            if (t_first != 0) {          // resume: the state the previous launch left
                carry = ctl->carry;
                adler_a = ctl->adler_a;
                adler_b = ctl->adler_b;
                pw = ctl->pw;
                lbit = ctl->lbit;
                wbase = 0;
            }
        }
        for (uint32_t t0 = t_first; t0 < t_stop; t0 += kTile) {
            const bool last_tile = closing && t0 + kTile >= L;
            const uint32_t n_tile = last_tile ? L - t0 : kTile;

            // ---------------- load: HBM -> shared, 128-bit cp.async where the vector is inside the stream
            __syncwarp();
            for (int k = lane; k < kInBytes / 16; k += 32) {
                const int64_t off = (int64_t)t0 - 32 + 16 * k;
                if (off >= 0 && off + 16 <= (int64_t)L) {
                    cp_async16(in_s + 16 * k, src + off);
                } else {
                    uint32_t w[4] = {0, 0, 0, 0};
                    if (off + 16 > 0 && off < (int64_t)L) {
#pragma unroll
                        for (int b = 0; b < 16; ++b) {
                            const int64_t q = off + b;
                            if (q >= 0 && q < (int64_t)L) w[b >> 2] |= (uint32_t)src[q] << (8 * (b & 3));
                        }
                    }
                    reinterpret_cast<uint4 *>(in_s)[k] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            T[vprev] = 0;                // drop the last chunk seen (previous tile's look-ahead / previous stream)
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncwarp();

            // ---------------- phase A: byte-equality masks ----------------------------------------
            uint32_t s1 = 0, s2 = 0;
            {
                // The lane mask of the lanes holding the same byte is built with a shared-memory
                // atomic OR into the table, not with MATCH.ANY: that instruction occupies the SM-wide
                // ADU pipe for ~2 cycles per DISTINCT value among the 32 lanes (~40 cycles on this
                // data, profiles/r01_ubench_match.txt) and made the whole kernel ADU-bound.
                // The atomic's return value holds the lanes of the same value that went before this one; only
                // the lanes BELOW it matter for the mask, and when the hardware took the lanes in ascending
                // order (checked every chunk) that is what it holds: no read-back of the entry.
                const uint32_t lane_bit = 1u << lane;
                // one chunk: c = -1 the last chunk of the previous tile (fills the table only), 0 .. kChunks - 1 the
                // tile, kChunks the look-ahead chunk (masks only).  Adler-32 partial sums (CSTATIC / CHECKSUM,
                // deflate.py:826-831, 884-897) ride along in the tile's chunks; bytes past the stream end are zero.
                auto chunk = [&](const int c, auto store, auto sums) {
                    const int i = 32 * c + lane;                 // tile-relative position
                    const uint32_t v = in_s[32 + i];
                    const uint32_t mprev = T[v];
                    __syncwarp();
                    T[vprev] = 0;
                    __syncwarp();
                    uint32_t mcur = atomicOr(&T[v], lane_bit);
                    __syncwarp();
                    if (__any_sync(HDLZ_FULL_MASK, (mcur >> lane) != 0u)) mcur = T[v];   // a higher lane went first
                    vprev = v;
                    if (decltype(store)::value) Rw[i + (i >> 5)] = __funnelshift_r(mprev, mcur, lane);   // bit k <=> distance 32 - k
                    if (decltype(sums)::value) {
                        s1 += v;
                        s2 += v * (n_tile - (uint32_t)i);
                    }
                };
                if (t0 > 0) chunk(-1, std::false_type(), std::false_type());
#pragma unroll 2
                for (int c = 0; c < kChunks; ++c) chunk(c, std::true_type(), std::true_type());
                chunk(kChunks, std::true_type(), std::false_type());
            }
            {
                // Adler-32 of the tile (deflate.py:826-831, 884-897), folded in now so that the sums die here
                const uint32_t S1 = __reduce_add_sync(HDLZ_FULL_MASK, s1);
                const uint32_t S2 = __reduce_add_sync(HDLZ_FULL_MASK, s2);
                adler_b = (adler_b + n_tile * adler_a + S2) % 65521u;
                adler_a = (adler_a + S1) % 65521u;
            }
            if (L - 2 < t0 + kTile + 32) {                   // R[q] = 0 for q >= L - 2
                __syncwarp();                                // other lanes wrote these words in phase A
                for (int i = (int)(L - 2 > t0 ? L - 2 - t0 : 0) + lane; i < kTile + 32; i += 32) Rw[i + (i >> 5)] = 0;
            }
            __syncwarp();

            // ---------------- phase B + P1: tokens and parse DP, segment walked from the back -------
            unsigned long long H = 0;
            {
                uint32_t win[17];
                const int rbase = 33 * lane;
#pragma unroll
                for (int k = 0; k < 9; ++k) win[8 + k] = Rw[rbase + 33 + k];
                __syncwarp();
                // a token that starts at j reaches into the next segment only when j + 10 >= 32: the
                // groups of the lower positions skip that case at compile time
                auto group = [&](const int grp, auto can_exit) {
                    const int j0 = grp * 8;
#pragma unroll
                    for (int k = 0; k < 8; ++k) win[k] = Rw[rbase + j0 + k];
                    const uint2 bv = *reinterpret_cast<const uint2 *>(in_s + 32 + 32 * lane + j0);
                    uint32_t tokv[8];
#pragma unroll
                    for (int k = 7; k >= 0; --k) {
                        const int j = j0 + k;
                        const uint32_t m3 = win[k] & win[k + 1] & win[k + 2];
                        const uint32_t f = find_msb(m3);                    // bit of the nearest distance, 0xFFFFFFFF: no match
                        uint32_t cbit = shl_clamp(1u, f);
                        cbit &= win[k + 3];
                        unsigned long long sum = cbit;
#pragma unroll
                        for (int e = 4; e < kMaxMatch; ++e) {
                            cbit &= win[k + e];
                            sum = add_wide(cbit, sum);
                        }
                        const uint32_t n = (uint32_t)shr64_clamp(sum, f);     // sum = n << f
                        const uint32_t x = ((k < 4 ? bv.x : bv.y) >> (8 * (k & 3))) & 255u;
                        uint32_t tk;
                        if constexpr (kMode == kModeFixed) {
                            const uint32_t mt = (m3 ? DC[f] : 0u) + (__brev(n + 1) >> 25) + (n << 26);
                            const uint32_t lt = LT[x];
                            tk = m3 ? mt : lt;
                        } else {
                            // a compact record, not the token: bits 24..29 = 4 * (length - 1) (what the parse needs),
                            // bit 8 = match, low bits the mask bit of the distance or the literal byte; P3 looks the
                            // codes up (or counts the symbols)
                            tk = m3 ? (n << 26) + ((8u << 24) | 0x100u) + f : x;
                        }
                        tokv[k] = tk;
                        const uint32_t ls = tk >> 24;                       // 4 * (length - 1)
                        const uint32_t look = (uint32_t)(H >> ls) & 15u;
                        if (decltype(can_exit)::value) {
                            const int ex = j + 1 + (int)(ls >> 2) - kSeg;
                            H = (H << 4) | (ex >= 0 ? (uint32_t)ex : look);
                        } else {
                            H = (H << 4) | look;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) Rw[rbase + j0 + k] = tokv[k];
#pragma unroll
                    for (int k = 8; k >= 0; --k) win[8 + k] = win[k];
                };
                group(3, std::true_type());
                group(2, std::true_type());
                for (int grp = 1; grp >= 0; --grp) group(grp, std::false_type());
            }
            __syncwarp();
            if (kMode == kModeFixed && last_tile) {            // positions past the end of the stream emit nothing
                for (int i = (int)n_tile + lane; i < kTile; i += 32) Rw[i + (i >> 5)] = 0;
                __syncwarp();
            }

            // ---------------- P2: entry skip count of every segment ----------------------------------
            uint32_t entry = 0;
            {
                // every lane publishes its 10-nibble map, then all lanes walk the 32 segments with
                // broadcast loads (one wavefront each) instead of two shuffles per step
                unsigned long long *Hs = reinterpret_cast<unsigned long long *>(ws.stage);    // input bytes are dead
                Hs[lane] = H;
                __syncwarp();
                if constexpr (kLong) {
                    // the tile's map carry-in -> carry-out: lanes 0..9 walk the segments, each from its own carry-in
                    uint32_t mc = lane < 10 ? (uint32_t)lane : 0u;
#pragma unroll 8
                    for (int s = 0; s < 32; ++s) mc = (uint32_t)(Hs[s] >> (4 * mc)) & 15u;
                    const uint32_t m_lo = __reduce_or_sync(HDLZ_FULL_MASK, lane < 8 ? mc << (4 * lane) : 0u);
                    const uint32_t m_hi = __reduce_or_sync(HDLZ_FULL_MASK, (lane == 8 || lane == 9) ? mc << (4 * (lane - 8)) : 0u);
                    const unsigned long long M = (unsigned long long)m_lo | ((unsigned long long)m_hi << 32);
                    const uint32_t c0 = __shfl_sync(HDLZ_FULL_MASK, mc, 0);
                    const bool constant = __all_sync(HDLZ_FULL_MASK, lane >= 10 || mc == c0);
                    uint32_t cin = 0;
                    if (lane == 0) {
                        const bool known = lt == 0 || constant;      // the carry-out does not depend on what comes in
                        if (lt == 0 && pc_t0 != 0) cin = ctl->carry;
                        if (known) atomicExch(&lb_map[lt], kLbResolved | (constant ? c0 : (uint32_t)(M >> (4 * cin)) & 15u));
                        else atomicExch(&lb_map[lt], kLbPartial | M);
                        if (lt != 0) {
                            // look back: compose the maps of the predecessors until one of them knows its carry-out
                            // (or the composition has become constant)
                            unsigned long long comp = 0x9876543210ull;     // identity
                            for (uint64_t j = lt - 1;; --j) {
                                const unsigned long long v = lb_poll(&lb_map[j]);
                                if ((v & kLbResolved) != 0ull) {
                                    cin = (uint32_t)(comp >> (4 * (v & 15ull))) & 15u;
                                    break;
                                }
                                unsigned long long nc = 0;                 // comp after M_j: e -> comp(M_j(e))
                                bool same = true;
                                for (int e = 0; e < 10; ++e) {
                                    const uint32_t x = (uint32_t)(comp >> (4 * ((v >> (4 * e)) & 15ull))) & 15u;
                                    nc |= (unsigned long long)x << (4 * e);
                                    same = same && x == (uint32_t)(nc & 15ull);
                                }
                                comp = nc;
                                if (same) {
                                    cin = (uint32_t)(comp & 15ull);
                                    break;
                                }
                            }
                            if (!known) {
                                __threadfence();
                                atomicExch(&lb_map[lt], kLbResolved | ((M >> (4 * cin)) & 15ull));
                            }
                        }
                    }
                    carry = __shfl_sync(HDLZ_FULL_MASK, cin, 0);
                }
                uint32_t cur = carry;
#pragma unroll 8
                for (int s = 0; s < 32; ++s) {
                    if (lane == s) entry = cur;
                    cur = (uint32_t)(Hs[s] >> (4 * cur)) & 15u;
                }
                carry = cur;
                __syncwarp();
            }

            // ---------------- P3: this lane's tokens -> its private bitstream --------------------------
            uint32_t nbits;
            if constexpr (kMode != kModeFixed) {
                // records -> codes of the tree (a match is two puts: length symbol, then distance code + extra bits;
                // DISTANCE, deflate.py:836-882), or -> counts
                uint32_t r = entry, fill = 0, wcnt = 0;
                unsigned long long acc = 0;
                const int rbase = 33 * lane;
                const int jend = (int)n_tile - 32 * lane;       // positions of the segment that exist (the last tile ends inside one)
#pragma unroll 4
                for (int j = 0; j < kSeg; ++j) {
                    const uint32_t rec = Rw[rbase + j];
                    const bool start = r == 0 && j < jend;
                    const bool is_match = (rec & 0x100u) != 0u;
                    const uint32_t n = ((rec >> 24) - 8u) >> 2;                 // extension of a match (length 3 + n)
                    const uint32_t i1 = is_match ? (uint32_t)kTreeLenBase + n : rec & 255u;
                    const uint32_t i2 = 256u + (rec & 31u);
                    if constexpr (kMode == kModeHist) {
                        if (start) {
                            atomicAdd(&LT[i1], 1u);
                            if (is_match) atomicAdd(&LT[i2], 1u);
                        }
                    } else {
                        uint32_t e1 = LT[i1];
                        uint32_t e2 = is_match ? LT[i2] : 0u;
                        if (start && ((e1 >> 16) == 0u || (is_match && (e2 >> 24) == 0u))) uncoded = 1;
                        if (!start) { e1 = 0; e2 = 0; }
                        acc |= (unsigned long long)(e1 & 0xFFFFu) << fill;
                        fill += e1 >> 16;
                        if (fill >= 32) {
                            priv[wcnt * 32 + lane] = (uint32_t)acc;
                            ++wcnt;
                            acc >>= 32;
                            fill -= 32;
                        }
                        acc |= (unsigned long long)(e2 & 0xFFFFFFu) << fill;
                        fill += e2 >> 24;
                        if (fill >= 32) {
                            priv[wcnt * 32 + lane] = (uint32_t)acc;
                            ++wcnt;
                            acc >>= 32;
                            fill -= 32;
                        }
                    }
                    r = r == 0 ? rec >> 26 : r - 1;
                }
                if (fill) priv[wcnt * 32 + lane] = (uint32_t)acc;
                nbits = 32 * wcnt + fill;
            } else {
                uint32_t r = entry, fill = 0, wcnt = 0;
                unsigned long long acc = 0;
                const int rbase = 33 * lane;
#pragma unroll 4
                for (int j = 0; j < kSeg; j += 2) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t tk = Rw[rbase + j + u];
                        const bool start = r == 0;
                        const uint32_t tke = start ? tk : 0u;
                        acc |= (unsigned long long)(tke & 0x7FFFu) << fill;
                        fill += (tke >> 16) & 15u;
                        r = start ? (tk >> 26) : r - 1;
                    }
                    if (fill >= 32) {                        // fill < 32 + 2 * 15 < 64 between checks
                        priv[wcnt * 32 + lane] = (uint32_t)acc;
                        ++wcnt;
                        acc >>= 32;
                        fill -= 32;
                    }
                }
                if (fill) priv[wcnt * 32 + lane] = (uint32_t)acc;     // wcnt <= kPrivWords - 1 here
                nbits = 32 * wcnt + fill;
            }
            if constexpr (kMode == kModeHist) continue;       // counted: nothing is written
            uint32_t incl = nbits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(HDLZ_FULL_MASK, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t tile_bits = __shfl_sync(HDLZ_FULL_MASK, incl, 31);
            if constexpr (kLong) {
                // the tile's place in the stream: bits of all earlier tiles (after the header bits of tile 0)
                unsigned long long b0 = 0;
                const bool tile_uncoded = kMode == kModeTree && __any_sync(HDLZ_FULL_MASK, uncoded != 0u);
                if (lane == 0) {
                    lb_adler[lt] = adler_a | (adler_b << 16);      // this tile's own sums (from a = 1, b = 0)
                    if (tile_uncoded) atomicOr(lb_flag, 1u);
                    __threadfence();
                    if (lt == 0) atomicExch(&lb_bits[0], kLbResolved | (32ull * wbase + lbit + tile_bits));
                    else atomicExch(&lb_bits[lt], kLbPartial | (unsigned long long)tile_bits);
                }
                if (lt != 0) {
                    // look back 32 tiles at a time (lane l reads tile end - 1 - l): bit counts are added up to and
                    // including the nearest tile that already knows where it ends.  A whole wave of tiles computes
                    // at once, so a one-entry-at-a-time walk would be thousands of dependent reads long.
                    for (uint64_t end = lt;; end -= 32) {
                        const bool in_range = end > (uint64_t)lane;
                        const unsigned long long v = in_range ? lb_poll(&lb_bits[end - 1 - lane]) : 0ull;
                        const uint32_t res = __ballot_sync(HDLZ_FULL_MASK, in_range && (v & kLbResolved) != 0ull);
                        const int stop = res ? __ffs(res) - 1 : 31;
                        unsigned long long x = (in_range && lane <= stop) ? (v & kLbValue) : 0ull;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(HDLZ_FULL_MASK, x, d);
                        b0 += x;
                        if (res) break;                               // tile 0 is always resolved: the walk ends there at the latest
                    }
                    if (lane == 0) {
                        __threadfence();
                        atomicExch(&lb_bits[lt], kLbResolved | (b0 + tile_bits));
                    }
                    pw = 0;
                    lbit = (uint32_t)(b0 & 31ull);
                    wbase = (uint32_t)(b0 >> 5);
                }
            }

            // the tokens are dead: their array becomes the tile's part of the output stream
            __syncwarp();
            for (int k = lane; k < kOutWords / 4; k += 32) reinterpret_cast<uint4 *>(outw)[k] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (lane == 0) outw[0] = pw;
            __syncwarp();

            // ---------------- merge: private streams -> the tile's stream at their bit offsets ----------
            {
                const uint32_t bp = lbit + incl - nbits;
                const uint32_t sh = bp & 31, w0 = bp >> 5;
                const uint32_t nsrc = (nbits + 31) >> 5;
                const uint32_t nwo = nbits ? (sh + nbits + 31) >> 5 : 0;
                const uint32_t nmax = __reduce_max_sync(HDLZ_FULL_MASK, nwo);
                const bool tail_partial = ((sh + nbits) & 31u) != 0;
                uint32_t prev = 0;
                for (uint32_t i = 0; i < nmax; ++i) {
                    const uint32_t cur = i < nsrc ? priv[i * 32 + lane] : 0u;
                    const uint32_t val = __funnelshift_l(prev, cur, sh);      // (cur << sh) | (prev >> (32 - sh))
                    prev = cur;
                    if (i < nwo) {
                        if ((i == 0 && sh != 0) || (i == nwo - 1 && tail_partial)) atomicOr(&outw[w0 + i], val);
                        else outw[w0 + i] = val;
                    }
                }
            }
            __syncwarp();

            // kLong: Adler-32 of everything up to and including this tile from the tiles' sums: (a1, b1, n1) then
            // (a2, b2, n2) give a = a1 + a2 - 1, b = b1 + b2 + n2 (a1 - 1)  (mod 65521).  Every earlier tile of the
            // launch is full.  Each lane folds a run of tiles, then the 32 runs and this tile are folded onto what the
            // earlier pieces left (1, 0 at the start of a stream).  Called by the last tile of a launch.
            auto fold_adler = [&]() {
                const uint64_t per = (lt + 31) / 32;
                const uint64_t j0 = per * lane < lt ? per * lane : lt, j1 = j0 + per < lt ? j0 + per : lt;
                uint32_t ra = 1, rb = 0;                       // 32-bit: b + b2 + 1024 * 65520 < 2^27
                unsigned long long rn = 0;
                for (uint64_t j = j0; j < j1; ++j) {
                    const uint32_t v = *reinterpret_cast<const volatile uint32_t *>(&lb_adler[j]);
                    const uint32_t a2 = v & 0xFFFFu, b2 = v >> 16;
                    rb = (rb + b2 + (uint32_t)kTile * ((ra + 65520u) % 65521u)) % 65521u;
                    ra = (ra + a2 + 65520u) % 65521u;
                    rn += kTile;
                }
                unsigned long long fa = 1, fb = 0;
                if (pc_t0 != 0) {
                    fa = ctl->adler_a;
                    fb = ctl->adler_b;
                }
                for (int l = 0; l < 32; ++l) {
                    const unsigned long long a2 = __shfl_sync(HDLZ_FULL_MASK, ra, l), b2 = __shfl_sync(HDLZ_FULL_MASK, rb, l);
                    const unsigned long long n2 = __shfl_sync(HDLZ_FULL_MASK, rn, l);
                    fb = (fb + b2 + (n2 % 65521ull) * (fa + 65520ull)) % 65521ull;
                    fa = (fa + a2 + 65520ull) % 65521ull;
                }
                const unsigned long long ta = adler_a, tb = adler_b;         // this tile's own
                adler_b = (uint32_t)((fb + tb + (unsigned long long)(n_tile % 65521u) * (fa + 65520ull)) % 65521ull);
                adler_a = (uint32_t)((fa + ta + 65520ull) % 65521ull);
            };
            // ---------------- flush ---------------------------------------------------------------------
            uint32_t total = lbit + tile_bits;
            if (kLong && !last_tile) {
                // whole words of the tile are its own; the first (if it starts inside one) and the last partial
                // word are shared with the neighbouring tiles: OR into the zeroed output
                const uint32_t nfull = total >> 5;
                for (uint32_t k = lane; k < nfull; k += 32) {
                    if (k == 0 && lbit != 0) atomicOr(&dst32[wbase], outw[0]);
                    else dst32[wbase + k] = outw[k];
                }
                if (lane == 0 && (total & 31u) != 0u) atomicOr(&dst32[wbase + nfull], outw[nfull]);
                if (ctl && !closing && lt + 1 == n_tiles) {      // the piece ends here: what the next one starts from
                    fold_adler();
                    if (lane == 0) {
                        ctl->carry = carry;
                        ctl->adler_a = adler_a;
                        ctl->adler_b = adler_b;
                        ctl->lbit = total & 31u;
                        ctl->out_words = wbase + nfull;          // the partial word after them stays in `out` for the host
                        ctl->t0 = t_stop;
                    }
                }
            } else if (!last_tile) {
                const uint32_t nfull = total >> 5;
                for (uint32_t k = lane; k < nfull; k += 32) dst32[wbase + k] = outw[k];
                pw = outw[nfull];
                wbase += nfull;
                lbit = total & 31;
            } else {
                if constexpr (kMode == kModeTree) {           // EOB with the tree's code for symbol 256
                    const uint32_t eob = tree->eob;
                    if (lane == 0) {
                        const unsigned long long v = (unsigned long long)(eob & 0xFFFFu) << (total & 31u);
                        outw[total >> 5] |= (uint32_t)v;
                        outw[(total >> 5) + 1] |= (uint32_t)(v >> 32);
                    }
                    __syncwarp();
                    total += eob >> 16;
                } else {
                    total += 7;                               // EOB: seven zero bits (deflate.py:772-779)
                }
                const uint32_t nbytes = (total + 7) >> 3;     // pad to a byte (deflate.py:784-787)
                uint32_t trailer = 4;
                if constexpr (kLong) fold_adler();
                if (lane == 0) {
                    uint8_t *ob = reinterpret_cast<uint8_t *>(outw);
                    if (container == HDLZ_CONTAINER_ZLIB) {
                        ob[nbytes + 0] = (uint8_t)(adler_b >> 8); // Adler-32 big-endian (deflate.py:788-814)
                        ob[nbytes + 1] = (uint8_t)(adler_b & 255);
                        ob[nbytes + 2] = (uint8_t)(adler_a >> 8);
                        ob[nbytes + 3] = (uint8_t)(adler_a & 255);
                    } else if (container == HDLZ_CONTAINER_GZIP) {
                        // CRC-32 (filled in by k_gzip_crc, which reads the input once more) and ISIZE
                        for (int b = 0; b < 4; ++b) {
                            ob[nbytes + b] = 0;
                            ob[nbytes + 4 + b] = (uint8_t)(L >> (8 * b));
                        }
                    }
                }
                if (container == HDLZ_CONTAINER_RAW) trailer = 0;
                else if (container == HDLZ_CONTAINER_GZIP) trailer = 8;
                __syncwarp();
                const uint32_t nwords = (nbytes + trailer + 3) >> 2;
                for (uint32_t k = lane; k < nwords; k += 32) {
                    if (kLong && k == 0 && lbit != 0) atomicOr(&dst32[wbase], outw[0]);      // shared with the tile before
                    else dst32[wbase + k] = outw[k];
                }
                bool no_code = kMode == kModeTree && __any_sync(HDLZ_FULL_MASK, uncoded != 0u);
                if (kLong && kMode == kModeTree) no_code = no_code || *reinterpret_cast<const volatile uint32_t *>(lb_flag) != 0u;
                if (lane == 0) {
                    out_len[ls] = no_code ? 0u : 4 * wbase + nbytes + trailer;
                    if (status) status[ls] = no_code ? HDLZ_ST_NO_CODE : HDLZ_OK;
                }
            }
        }
        if (kStream) {
            __syncwarp();
            if (lane == 0) {
                if (!closing) {
                    ctl->t0 = t_stop;
                    ctl->carry = carry;
                    ctl->adler_a = adler_a;
                    ctl->adler_b = adler_b;
                    ctl->pw = pw;
                    ctl->lbit = lbit;
                }
                ctl->out_words = wbase;
            }
        }
        sid = n_warps + __shfl_sync(HDLZ_FULL_MASK, next_ticket, 0);
    }
    if constexpr (kMode == kModeHist) {
        __syncthreads();
        for (int i = threadIdx.x; i < kTreeHistWords; i += kWarpsPerCta * 32)
            if (LT[i]) atomicAdd(&hist[i], (unsigned long long)LT[i]);
    }
}

}  // namespace

// One launch of the stream kernel (hdlz_stream_feed / hdlz_stream_finish, hdlz_api.cu): a single warp.
int launch_compress_stream(hdlz_ctx *ctx, const uint8_t *d_in_virtual, uint32_t received, uint8_t *d_out, uint32_t *d_out_len,
                           uint32_t *d_status, StreamCtl *d_ctl, unsigned long long *d_queue, cudaStream_t s)
{
    if (!ctx->stream_attr_set) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, true, kModeFixed>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kModeFixed>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, true, kModeFixed>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kModeFixed>()));
        ctx->stream_attr_set = true;
    }
    if (ctx->max_match == 5)
        k_compress<5, true, kModeFixed><<<1, kWarpsPerCta * 32, smem_bytes<kModeFixed>(), s>>>(
            d_in_virtual, 0, nullptr, received, d_out, 0, d_out_len, d_status, 1, d_queue, ctx->container, d_ctl, nullptr, nullptr);
    else
        k_compress<10, true, kModeFixed><<<1, kWarpsPerCta * 32, smem_bytes<kModeFixed>(), s>>>(
            d_in_virtual, 0, nullptr, received, d_out, 0, d_out_len, d_status, 1, d_queue, ctx->container, d_ctl, nullptr, nullptr);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

// queue head of one launch of the persistent grid: the next of kQueueSlots slots, zeroed on the launch's own
// stream.  A slot comes round again after kQueueSlots compress launches of this context, far more than can be
// in flight on its streams at once.
static int next_queue(hdlz_ctx *ctx, unsigned long long **queue, cudaStream_t s)
{
    constexpr unsigned kQueueSlots = 4096;
    if (!ctx->d_queue) HDLZ_CUDA(cudaMalloc((void **)&ctx->d_queue, kQueueSlots * sizeof(unsigned long long)));
    *queue = ctx->d_queue + (ctx->queue_seq++ & (kQueueSlots - 1));
    HDLZ_CUDA(cudaMemsetAsync(*queue, 0, sizeof(unsigned long long), s));
    return HDLZ_SUCCESS;
}

template <int kMode>
static int launch_mode(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len, uint32_t uniform_len,
                       uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len, uint32_t *d_status, uint64_t n,
                       unsigned long long *d_hist, cudaStream_t s)
{
    bool &attr = kMode == kModeFixed ? ctx->compress_attr_set : ctx->tree_attr_set;
    if (!attr) {
        constexpr int kA = kMode == kModeFixed ? kModeFixed : kModeTree, kB = kMode == kModeFixed ? kModeFixed : kModeHist;
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kA>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kA>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kA>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kA>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kB>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kB>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr = true;
    }
    uint64_t blocks = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint64_t resident = (uint64_t)ctx->sm_count * kCtasPerSm;      // persistent: a multiple of the SM count
    if (blocks > resident) blocks = resident;
    unsigned long long *queue = nullptr;
    int rc = next_queue(ctx, &queue, s);
    if (rc) return rc;
    if (ctx->max_match == 5)
        k_compress<5, false, kMode><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kMode>(), s>>>(
            d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n, queue, ctx->container, nullptr,
            ctx->d_tree, d_hist);
    else
        k_compress<10, false, kMode><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kMode>(), s>>>(
            d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n, queue, ctx->container, nullptr,
            ctx->d_tree, d_hist);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

// One long stream over the whole grid (k_compress<.., kLong = true>): hdlz_compress_stream for inputs of many
// tiles.  d_out must hold compress_bound(len) bytes; it is zeroed here (tiles OR their border words into it).
template <int kMode>
static int launch_long_mode(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, uint32_t len, uint8_t *d_out, uint64_t out_stride,
                            uint32_t *d_out_len, uint32_t *d_status, uint64_t n, cudaStream_t s)
{
    bool &attr = kMode == kModeFixed ? ctx->long_attr_set : ctx->long_tree_attr_set;
    if (!attr) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kMode, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kMode>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kMode, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kMode, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kMode>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kMode, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr = true;
    }
    const uint64_t n_tiles = (((uint64_t)len + kTile - 1) / kTile) * n;
    const size_t lb_bytes = n_tiles * (2 * sizeof(unsigned long long) + sizeof(uint32_t)) + n * sizeof(uint32_t) + 16;
    int rc = grow_device((void **)&ctx->d_long, &ctx->d_long_cap, lb_bytes);
    if (rc) return rc;
    HDLZ_CUDA(cudaMemsetAsync(ctx->d_long, 0, lb_bytes, s));
    HDLZ_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n * out_stride, s));
    unsigned long long *queue = nullptr;
    if ((rc = next_queue(ctx, &queue, s))) return rc;
    uint64_t blocks = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint64_t resident = (uint64_t)ctx->sm_count * kCtasPerSm;
    if (blocks > resident) blocks = resident;
    unsigned long long *lbuf = reinterpret_cast<unsigned long long *>(ctx->d_long);
    if (ctx->max_match == 5)
        k_compress<5, false, kMode, true><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kMode>(), s>>>(
            d_in, in_stride, nullptr, len, d_out, out_stride, d_out_len, d_status, n, queue, ctx->container, nullptr, ctx->d_tree, nullptr, lbuf);
    else
        k_compress<10, false, kMode, true><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kMode>(), s>>>(
            d_in, in_stride, nullptr, len, d_out, out_stride, d_out_len, d_status, n, queue, ctx->container, nullptr, ctx->d_tree, nullptr, lbuf);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

// One piece of a stream fed in pieces, many tiles of it: the long-stream kernel with the state in *d_ctl
// (hdlz_cstream_feed / _finish).  d_out (out_bytes, zeroed here) receives this launch's words.
int launch_compress_piece(hdlz_ctx *ctx, const uint8_t *d_in_virtual, uint32_t received, uint64_t n_tiles, uint8_t *d_out,
                          uint64_t out_bytes, uint32_t *d_out_len, uint32_t *d_status, StreamCtl *d_ctl, cudaStream_t s)
{
    if (!ctx->long_attr_set) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kModeFixed, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kModeFixed>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<10, false, kModeFixed, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kModeFixed, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<kModeFixed>()));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress<5, false, kModeFixed, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->long_attr_set = true;
    }
    const size_t lb_bytes = n_tiles * (2 * sizeof(unsigned long long) + sizeof(uint32_t)) + sizeof(uint32_t) + 16;
    int rc = grow_device((void **)&ctx->d_long, &ctx->d_long_cap, lb_bytes);
    if (rc) return rc;
    HDLZ_CUDA(cudaMemsetAsync(ctx->d_long, 0, lb_bytes, s));
    HDLZ_CUDA(cudaMemsetAsync(d_out, 0, out_bytes, s));
    unsigned long long *queue = nullptr;
    if ((rc = next_queue(ctx, &queue, s))) return rc;
    uint64_t blocks = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint64_t resident = (uint64_t)ctx->sm_count * kCtasPerSm;
    if (blocks > resident) blocks = resident;
    unsigned long long *lbuf = reinterpret_cast<unsigned long long *>(ctx->d_long);
    if (ctx->max_match == 5)
        k_compress<5, false, kModeFixed, true><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kModeFixed>(), s>>>(
            d_in_virtual, 0, nullptr, received, d_out, 0, d_out_len, d_status, 1, queue, ctx->container, d_ctl, nullptr, nullptr, lbuf);
    else
        k_compress<10, false, kModeFixed, true><<<(unsigned)blocks, kWarpsPerCta * 32, smem_bytes<kModeFixed>(), s>>>(
            d_in_virtual, 0, nullptr, received, d_out, 0, d_out_len, d_status, 1, queue, ctx->container, d_ctl, nullptr, nullptr, lbuf);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

int launch_compress_long(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, uint32_t len, uint8_t *d_out, uint64_t out_stride,
                         uint32_t *d_out_len, uint32_t *d_status, uint64_t n, cudaStream_t s)
{
    if (ctx->tree_set) {
        const int rc = refresh_tree(ctx);
        if (rc) return rc;
        return launch_long_mode<kModeTree>(ctx, d_in, in_stride, len, d_out, out_stride, d_out_len, d_status, n, s);
    }
    return launch_long_mode<kModeFixed>(ctx, d_in, in_stride, len, d_out, out_stride, d_out_len, d_status, n, s);
}

// hdlz_train_tree: the symbols of the reference's parse over the batch, counted into d_hist[kTreeHistWords]
int launch_compress_hist(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len, uint32_t uniform_len,
                         uint64_t n, unsigned long long *d_hist, cudaStream_t s)
{
    return launch_mode<kModeHist>(ctx, d_in, in_stride, d_in_len, uniform_len, nullptr, 0, nullptr, nullptr, n, d_hist, s);
}

int launch_compress(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                    uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                    uint32_t *d_status, uint64_t n, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    int rc;
    if (ctx->window == 256) {
        // the reference's non-FAST configuration (CWINDOW = 256): hdlz_compress_wide.cu
        if (ctx->tree_set) return set_error(HDLZ_ERR_INVALID, "a tree (hdlz_set_tree) needs the FAST compressor (CWINDOW = 32)");
        unsigned long long *queue = nullptr;
        if ((rc = next_queue(ctx, &queue, s))) return rc;
        rc = launch_compress_wide(ctx, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n, queue, s);
    } else if (ctx->tree_set) {
        if ((rc = refresh_tree(ctx))) return rc;
        rc = launch_mode<kModeTree>(ctx, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n, nullptr, s);
    } else {
        rc = launch_mode<kModeFixed>(ctx, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, d_status, n, nullptr, s);
    }
    if (rc) return rc;
    if (ctx->container == HDLZ_CONTAINER_GZIP)
        return launch_gzip_trailers(ctx, d_in, in_stride, d_in_len, uniform_len, d_out, out_stride, d_out_len, n, s);
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
