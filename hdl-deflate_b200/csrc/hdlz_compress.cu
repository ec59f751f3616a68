// hdlz_compress.cu — static-tree deflate compressor, bit-exact with the reference's
// FAST + MATCH10 / CWINDOW=32 engine (deflate.py states CSTATIC, SEARCH, SEARCHF,
// DISTANCE, CHECKSUM; :734-1016), re-designed for sm_100a.  Not a port of the FSM:
//
//   one WARP per stream, tiles of 2048 input positions, everything between the
//   HBM read of the input and the HBM write of the stream lives in shared memory
//   and registers.
//
//   phase A  lane <-> position (32 positions per step)
//            R[p] = 32-bit mask, bit (32-d) <=> x[p-d] == x[p], d = 1..32.
//            Inside a 32-chunk the equalities come from ONE match.any; against the
//            previous chunk from a 256-entry per-warp table T[value] = lane mask of
//            that value in the previous chunk; one funnel shift joins both halves.
//            This replaces the 32 `matcher3` comparators + cwindow shift register
//            (deflate.py:407-421, 442-453).  R[q] = 0 for q >= L-2 encodes all the
//            `di < isize - k` guards of SEARCH/SEARCHF (deflate.py:913-952, 975-977).
//            Adler-32 partial sums ride along (CSTATIC/CHECKSUM, :826-831, :884-897).
//   phase B  lane <-> segment of 64 consecutive positions, walked from the back
//            M3 = R[p] & R[p+1] & R[p+2]  -> 3-byte match at every distance at once;
//            nearest distance = clz (lowest `si` first, deflate.py:982-988);
//            length = 3 + leading run of that bit through R[p+3..p+9] (SEARCHF);
//            token bits from two small LUTs (fixed Huffman, DISTANCE :836-882).
//            The same backward walk runs the parse DP: h[j] = skip count left for
//            the next segment if a token starts at j (10-nibble shift register).
//   phase P2 resolve the entry skip count of each of the 32 segments (greedy parse
//            `di += match` / `di += 1`, deflate.py:960, 1008) with 32 shuffles.
//   phase P3 forward walk: mark token starts, sum their bit lengths, warp scan,
//            then OR the codes into the staged bitstream at their bit offsets
//            (put/do_flush, deflate.py:535-567).
//   flush    whole words of the staged stream go to HBM with coalesced stores; the
//            last tile appends EOB, pad and Adler-32 (deflate.py:771-814).
//
// Algorithmic HBM traffic per stream: L bytes read + C bytes written (C = stream size).

#include "hdlz_common.cuh"

namespace hdlz {
namespace {

constexpr int kTile = 2048;            // positions per tile
constexpr int kSeg = 64;               // positions per lane in the segment phases
constexpr int kWarpsPerCta = 4;
constexpr int kInBytes = 32 + kTile + 32;          // history | tile | look-ahead chunk
constexpr int kRWords = (kTile + 32) + (kTile + 32) / kSeg + 1;   // padded: idx = i + i/64
constexpr int kOutWords = 592;         // 31 carry bits + 2048*9 + EOB + Adler, rounded up
constexpr int kStageBytes = kOutWords * 4 > kInBytes ? kOutWords * 4 : kInBytes;

struct __align__(16) WarpSmem {
    uint8_t stage[(kStageBytes + 15) / 16 * 16];   // input tile, later the staged output words
    uint32_t R[(kRWords + 3) / 4 * 4];             // masks, overwritten in place by tokens
    uint32_t T[256];                               // value -> lane mask of the previous chunk
};

constexpr int kLutWords = 512;   // MT[256] match tokens, LT[256] literal tokens
constexpr size_t kSmemBytes = kLutWords * 4 + sizeof(WarpSmem) * kWarpsPerCta;

// token word: bits 0..14 code (LSB-first), 16..19 bit count, 24..27 length in positions
__device__ __forceinline__ uint32_t rev_n(uint32_t v, int n) { return __brev(v) >> (32 - n); }

__device__ __forceinline__ uint32_t match_token_entry(int n, int cl)
{
    // length symbol 257 + n (length 3 + n): 7-bit code n + 1 (RFC 1951 3.2.6; lencode = mlength + 254)
    uint32_t code = rev_n((uint32_t)(n + 1), 7);
    // distance d = cl + 1 -> code c, extra bits (CopyDistance / ExtraDistanceBits, deflate.py:106-110)
    uint32_t e = (uint32_t)cl, c, eb, extra;
    if (e < 4) {
        c = e; eb = 0; extra = 0;
    } else {
        uint32_t msb = 31 - __clz(e);
        eb = msb - 1;
        c = 2 * msb + ((e >> eb) & 1);
        extra = e & ((1u << eb) - 1);
    }
    uint32_t dc = rev_n(c, 5) | (extra << 5);
    return (code | (dc << 7)) | ((12 + eb) << 16);
}

__device__ __forceinline__ uint32_t literal_token_entry(uint32_t x)
{
    // fixed Huffman literal codes, bit-reversed (== out_codes[x], deflate.py:112-149)
    return x < 144 ? (rev_n(0x30 + x, 8) | (8u << 16)) : (rev_n(0x100 + x, 9) | (9u << 16));
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_compress(const uint8_t *__restrict__ in, uint64_t in_stride, const uint32_t *__restrict__ in_len,
           uint32_t uniform_len, uint8_t *__restrict__ out, uint64_t out_stride,
           uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams)
{
    extern __shared__ uint4 smem_raw[];
    uint32_t *MT = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *LT = MT + 256;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(MT + kLutWords)[warp];

    for (int i = threadIdx.x; i < 256; i += kWarpsPerCta * 32) {
        MT[i] = match_token_entry(i >> 5, i & 31);
        LT[i] = literal_token_entry((uint32_t)i);
    }
    __syncthreads();

    const uint64_t sid = (uint64_t)blockIdx.x * kWarpsPerCta + warp;
    if (sid >= n_streams) return;

    const uint32_t L = in_len ? in_len[sid] : uniform_len;
    const uint8_t *src = in + sid * in_stride;
    uint8_t *dst = out + sid * out_stride;
    if (L < HDLZ_MIN_INPUT || (uint64_t)compress_bound(L) > out_stride) {
        if (lane == 0) {
            out_len[sid] = 0;
            if (status) status[sid] = L < HDLZ_MIN_INPUT ? HDLZ_ST_SHORT_INPUT : HDLZ_ST_OUT_OVERFLOW;
        }
        return;
    }

    uint8_t *in_s = ws.stage;                                  // byte i <-> position t0 - 32 + i
    uint32_t *outw = reinterpret_cast<uint32_t *>(ws.stage);   // staged output words (after phase B)
    uint32_t *Rw = ws.R;
    uint32_t *T = ws.T;

    for (int i = lane; i < 256; i += 32) T[i] = 0;

    uint32_t carry = 0;              // positions of the next tile still covered by the last token
    uint32_t adler_a = 1, adler_b = 0;
    uint32_t pw = 0x78u | (0x9Cu << 8) | (3u << 16);   // partial output word: header + BFINAL/BTYPE=01
    uint32_t lbit = 19;              // valid bits in pw
    uint32_t wbase = 0;              // 32-bit words of the stream already written to HBM
    uint32_t vprev = 0;
    uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst);

    for (uint32_t t0 = 0; t0 < L; t0 += kTile) {
        const bool last_tile = t0 + kTile >= L;
        const uint32_t n_tile = last_tile ? L - t0 : kTile;

        // ---------------- load: HBM -> shared, 128-bit when fully inside the stream ----------
        __syncwarp();
        for (int k = lane; k < kInBytes / 16; k += 32) {
            const int64_t off = (int64_t)t0 - 32 + 16 * k;
            uint4 v;
            if (off >= 0 && off + 16 <= (int64_t)L) {
                v = *reinterpret_cast<const uint4 *>(src + off);
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int64_t q = off + b;
                    if (q >= 0 && q < (int64_t)L) w[b >> 2] |= (uint32_t)src[q] << (8 * (b & 3));
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            reinterpret_cast<uint4 *>(in_s)[k] = v;
        }
        if (t0 > 0) T[vprev] = 0;    // drop the look-ahead chunk of the previous tile from the table
        __syncwarp();

        // ---------------- phase A: byte-equality masks ----------------------------------------
        uint32_t s1 = 0, s2 = 0;
        for (int c = (t0 > 0 ? -1 : 0); c <= kTile / 32; ++c) {
            const int i = 32 * c + lane;                 // tile-relative position
            const uint32_t v = in_s[32 + i];
            const uint32_t mcur = __match_any_sync(HDLZ_FULL_MASK, v);
            const uint32_t mprev = T[v];
            __syncwarp();
            if (c > (t0 > 0 ? -1 : 0)) T[vprev] = 0;
            __syncwarp();
            T[v] = mcur;
            __syncwarp();
            vprev = v;
            if (c >= 0) {
                uint32_t r = __funnelshift_r(mprev, mcur, lane);   // bit k <=> distance 32 - k
                const uint32_t q = t0 + (uint32_t)i;
                if (q + 2 >= L) r = 0;
                Rw[i + (i >> 6)] = r;
                if (c < kTile / 32 && q < L) {
                    s1 += v;
                    s2 += v * (n_tile - (uint32_t)i);
                }
            }
        }
        __syncwarp();

        // ---------------- phase B + P1: tokens and parse DP, segment walked from the back -------
        uint64_t H = 0;
        {
            uint32_t win[25];
            const int rbase = 65 * lane;
#pragma unroll
            for (int k = 0; k < 9; ++k) win[16 + k] = Rw[rbase + 65 + k];
            __syncwarp();
            for (int grp = 3; grp >= 0; --grp) {
                const int j0 = grp * 16;
#pragma unroll
                for (int k = 0; k < 16; ++k) win[k] = Rw[rbase + j0 + k];
                const uint4 bv = *reinterpret_cast<const uint4 *>(in_s + 32 + 64 * lane + j0);
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
                uint32_t tokv[16];
#pragma unroll
                for (int k = 15; k >= 0; --k) {
                    const int j = j0 + k;
                    const uint32_t m3 = win[k] & win[k + 1] & win[k + 2];
                    const uint32_t cl = __clz(m3) & 31;
                    uint32_t cbit = m3 ? (0x80000000u >> cl) : 0u;
                    uint32_t n = 0;
#pragma unroll
                    for (int e = 3; e < 10; ++e) {
                        cbit &= win[k + e];
                        n += cbit ? 1u : 0u;
                    }
                    const uint32_t x = (bw[k >> 2] >> (8 * (k & 3))) & 255u;
                    const uint32_t mt = MT[n * 32 + cl];
                    const uint32_t lt = LT[x];
                    uint32_t ln = m3 ? 3 + n : 1;
                    uint32_t tk = (m3 ? mt : lt) | (ln << 24);
                    if (t0 + 64 * lane + j >= L) { tk = 1u << 24; ln = 1; }
                    tokv[k] = tk;
                    const int ex = j + (int)ln - kSeg;
                    const uint32_t hn = ex >= 0 ? (uint32_t)ex : ((uint32_t)(H >> (4 * (ln - 1))) & 15u);
                    H = (H << 4) | hn;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) Rw[rbase + j0 + k] = tokv[k];
#pragma unroll
                for (int k = 0; k < 9; ++k) win[16 + k] = win[k];
            }
        }
        __syncwarp();

        // ---------------- P2: entry skip count of every segment ----------------------------------
        uint32_t entry = 0;
        {
            const uint32_t hlo = (uint32_t)H, hhi = (uint32_t)(H >> 32) & 0xFFu;
            uint32_t cur = carry;
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                if (lane == s) entry = cur;
                const uint32_t lo = __shfl_sync(HDLZ_FULL_MASK, hlo, s);
                const uint32_t hi = __shfl_sync(HDLZ_FULL_MASK, hhi, s);
                cur = (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (4 * cur)) & 15u;
            }
            carry = cur;
        }

        // stage buffer changes role: input bytes are dead, zero it for the output words
        for (int k = lane; k < kOutWords / 4; k += 32) reinterpret_cast<uint4 *>(outw)[k] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        if (lane == 0) outw[0] = pw;

        // ---------------- P3a: bits emitted by this lane's segment --------------------------------
        uint32_t bits = 0;
        {
            uint32_t r = entry;
            const int rbase = 65 * lane;
#pragma unroll 8
            for (int j = 0; j < kSeg; ++j) {
                const uint32_t tk = Rw[rbase + j];
                const bool start = r == 0;
                bits += start ? ((tk >> 16) & 15u) : 0u;
                r = start ? (tk >> 24) - 1 : r - 1;
            }
        }
        uint32_t incl = bits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(HDLZ_FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t tile_bits = __shfl_sync(HDLZ_FULL_MASK, incl, 31);
        __syncwarp();

        // ---------------- P3b: OR the codes into the staged stream ---------------------------------
        {
            uint32_t bp = lbit + incl - bits;
            uint32_t w = bp >> 5;
            uint32_t fill = bp & 31;
            uint64_t acc = 0;
            uint32_t r = entry;
            const int rbase = 65 * lane;
#pragma unroll 4
            for (int j = 0; j < kSeg; ++j) {
                const uint32_t tk = Rw[rbase + j];
                if (r == 0) {
                    acc |= (uint64_t)(tk & 0x7FFFu) << fill;
                    fill += (tk >> 16) & 15u;
                    r = (tk >> 24) - 1;
                    if (fill >= 32) {
                        atomicOr(&outw[w], (uint32_t)acc);
                        ++w;
                        acc >>= 32;
                        fill -= 32;
                    }
                } else {
                    --r;
                }
            }
            if (fill) atomicOr(&outw[w], (uint32_t)acc);
        }
        __syncwarp();

        // ---------------- Adler-32 of the tile -------------------------------------------------------
        {
            const uint32_t S1 = __reduce_add_sync(HDLZ_FULL_MASK, s1);
            const uint32_t S2 = __reduce_add_sync(HDLZ_FULL_MASK, s2);
            adler_b = (adler_b + n_tile * adler_a + S2) % 65521u;
            adler_a = (adler_a + S1) % 65521u;
        }

        // ---------------- flush ---------------------------------------------------------------------
        uint32_t total = lbit + tile_bits;
        if (!last_tile) {
            const uint32_t nfull = total >> 5;
            for (uint32_t k = lane; k < nfull; k += 32) dst32[wbase + k] = outw[k];
            pw = outw[nfull];
            wbase += nfull;
            lbit = total & 31;
        } else {
            total += 7;                                   // EOB: seven zero bits (deflate.py:772-779)
            const uint32_t nbytes = (total + 7) >> 3;     // pad to a byte (deflate.py:784-787)
            if (lane == 0) {
                uint8_t *ob = reinterpret_cast<uint8_t *>(outw);
                ob[nbytes + 0] = (uint8_t)(adler_b >> 8); // Adler-32 big-endian (deflate.py:788-814)
                ob[nbytes + 1] = (uint8_t)(adler_b & 255);
                ob[nbytes + 2] = (uint8_t)(adler_a >> 8);
                ob[nbytes + 3] = (uint8_t)(adler_a & 255);
            }
            __syncwarp();
            const uint32_t nwords = (nbytes + 4 + 3) >> 2;
            for (uint32_t k = lane; k < nwords; k += 32) dst32[wbase + k] = outw[k];
            if (lane == 0) {
                out_len[sid] = 4 * wbase + nbytes + 4;
                if (status) status[sid] = HDLZ_OK;
            }
        }
    }
}

}  // namespace

int launch_compress(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                    uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                    uint32_t *d_status, uint64_t n, cudaStream_t s)
{
    static bool attr_set[64] = {false};
    if (n == 0) return HDLZ_SUCCESS;
    if (!attr_set[ctx->device & 63]) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        attr_set[ctx->device & 63] = true;
    }
    const uint64_t blocks = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    if (blocks > 0x7FFFFFFFull) return set_error(HDLZ_ERR_INVALID, "too many streams for one launch");
    k_compress<<<(unsigned)blocks, kWarpsPerCta * 32, kSmemBytes, s>>>(d_in, in_stride, d_in_len, uniform_len, d_out,
                                                                        out_stride, d_out_len, d_status, n);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
