// hdlz_compress_wide.cu — the reference's non-FAST compressor configuration (deflate.py:36-37, 56-59:
// FAST = False => CWINDOW = 256), bit-exact with it: SEARCH walks `cur_search` down from di - 1 while it stays
// within the window (:996-1016) — the NEAREST 3-byte match within 256 bytes wins — SEARCH10 grows it to at most
// 10 (or 5) bytes under the same `di < isize - k` guards (:1018-1062), and DISTANCE emits distance codes up to
// 15 whose 5 + 6 bits leave the FSM in two puts (`outcarry`, :875-880; the same bits in the stream).
//
// A second mode, not the hot path of the BASELINE configs: same structure as hdlz_compress.cu from the token
// array on (parse DP, entry resolution, lane-private bitstreams, merge, flush), but the tokens come from a
// direct search: 24-bit trigram words of the tile and its 256 bytes of history sit in shared memory, the 32
// lanes of a chunk walk the distances upwards together until each has found its nearest match.  Token words
// are wider here (18 code bits: 7 + 5 + 6).
//
// Algorithmic HBM traffic per stream: L bytes read + C bytes written.

#include "hdlz_common.cuh"

namespace hdlz {
namespace {

constexpr int kTile = 1024;
constexpr int kSeg = 32;
constexpr int kChunks = kTile / 32;
constexpr int kWarpsPerCta = 4;
constexpr int kCtasPerSm = 4;
constexpr int kHist = 256;                                  // CWINDOW of the non-FAST engine
constexpr int kInBytes = kHist + kTile + 32;                // history | tile | look-ahead
constexpr int kPrivWords = ((kSeg - 1) * 9 + 18 + 31) / 32; // 31 nine-bit literals + one 18-bit match token = 10 words
constexpr int kStageBytes = kInBytes;                       // 1312 >= kPrivWords * 128: input tile, later the private streams
constexpr int kRWords = kTile + kTile / kSeg;               // idx = i + i / 32
constexpr int kOutWords = 296;                              // 31 carry bits + 1024 * 9 + EOB + Adler, rounded up
static_assert(kStageBytes >= kPrivWords * 32 * 4 && kStageBytes % 16 == 0, "stage buffer");

struct __align__(16) WarpSmemW {
    uint8_t stage[kStageBytes];                             // byte i <-> position t0 - kHist + i
    uint32_t R[(kRWords + 4) / 4 * 4];                      // tokens; after P3 the tile's part of the output stream
    uint32_t W[kHist + kTile];                              // trigram x[q] | x[q+1] << 8 | x[q+2] << 16 of position t0 - kHist + i
};
constexpr size_t kSmemBytes = sizeof(WarpSmemW) * kWarpsPerCta;

// token word: bits 0..17 code (LSB-first), 18..22 bit count, 24..29 = 4 * (length in positions - 1)
__device__ __forceinline__ uint32_t rev_n(uint32_t v, int n) { return __brev(v) >> (32 - n); }

__device__ __forceinline__ uint32_t literal_token(uint32_t x)
{
    // fixed Huffman literal codes, bit-reversed (== out_codes[x], deflate.py:112-149)
    return x < 144 ? (rev_n(0x30 + x, 8) | (8u << 18)) : (rev_n(0x100 + x, 9) | (9u << 18));
}

__device__ __forceinline__ uint32_t match_token(uint32_t d, uint32_t m)
{
    // length symbol 254 + m: 7-bit code m - 2, no extra bits (deflate.py:845-850); distance code c with
    // ExtraDistanceBits[c // 2] extra bits (CopyDistance / ExtraDistanceBits, deflate.py:106-110, 858-874)
    const uint32_t e = d - 1;
    uint32_t c, eb, extra;
    if (e < 4) {
        c = e; eb = 0; extra = 0;
    } else {
        const uint32_t msb = 31 - __clz(e);
        eb = msb - 1;
        c = 2 * msb + ((e >> eb) & 1);
        extra = e & ((1u << eb) - 1);
    }
    const uint32_t dc = rev_n(c, 5) | (extra << 5);
    return rev_n(m - 2, 7) | (dc << 7) | ((12 + eb) << 18) | ((4 * (m - 1)) << 24);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

template <int kMaxMatch>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm)
k_compress_wide(const uint8_t *__restrict__ in, uint64_t in_stride, const uint32_t *__restrict__ in_len,
                uint32_t uniform_len, uint8_t *__restrict__ out, uint64_t out_stride,
                uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams,
                unsigned long long *queue, uint32_t container)
{
    extern __shared__ uint4 smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmemW &ws = reinterpret_cast<WarpSmemW *>(smem_raw)[warp];
    uint8_t *in_s = ws.stage;
    uint32_t *priv = reinterpret_cast<uint32_t *>(ws.stage);   // word k of lane l at priv[k * 32 + l]
    uint32_t *Rw = ws.R;
    uint32_t *Wt = ws.W;
    uint32_t *outw = ws.R;

    const uint64_t n_warps = (uint64_t)gridDim.x * kWarpsPerCta;
    for (uint64_t sid = (uint64_t)blockIdx.x * kWarpsPerCta + warp; sid < n_streams;) {
        unsigned long long next_ticket = 0;
        if (lane == 0) next_ticket = atomicAdd(queue, 1ull);
        const uint32_t L = in_len ? in_len[sid] : uniform_len;
        const uint8_t *src = in + sid * in_stride;
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(out + sid * out_stride);
        if (L < HDLZ_MIN_INPUT || (uint64_t)compress_bound(L, container) > out_stride) {
            if (lane == 0) {
                out_len[sid] = 0;
                if (status) status[sid] = L < HDLZ_MIN_INPUT ? HDLZ_ST_SHORT_INPUT : HDLZ_ST_OUT_OVERFLOW;
            }
            sid = n_warps + __shfl_sync(HDLZ_FULL_MASK, next_ticket, 0);
            continue;
        }

        uint32_t carry = 0;
        uint32_t adler_a = 1, adler_b = 0;
        uint32_t pw = 0x78u | (0x9Cu << 8) | (3u << 16);     // 78 9C, BFINAL = 1 / BTYPE = 01 (deflate.py:753-761)
        uint32_t lbit = 19, wbase = 0;
        if (container == HDLZ_CONTAINER_RAW) {
            pw = 3u;
            lbit = 3;
        } else if (container == HDLZ_CONTAINER_GZIP) {
            if (lane == 0) {
                dst32[0] = 0x00088B1Fu;
                dst32[1] = 0u;
            }
            pw = (0xFFu << 8) | (3u << 16);
            wbase = 2;
        }

        for (uint32_t t0 = 0; t0 < L; t0 += kTile) {
            const bool last_tile = t0 + kTile >= L;
            const uint32_t n_tile = last_tile ? L - t0 : kTile;

            // ---------------- load: history | tile | look-ahead, zero outside the stream
            __syncwarp();
            for (int k = lane; k < kInBytes / 16; k += 32) {
                const int64_t off = (int64_t)t0 - kHist + 16 * k;
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
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            for (int i = lane; i < kHist + kTile; i += 32)
                Wt[i] = (uint32_t)in_s[i] | ((uint32_t)in_s[i + 1] << 8) | ((uint32_t)in_s[i + 2] << 16);
            __syncwarp();

            // ---------------- SEARCH / SEARCH10: nearest 3-byte match within 256 bytes, grown to kMaxMatch
            uint32_t s1 = 0, s2 = 0;
            for (int c = 0; c < kChunks; ++c) {
                const int i = 32 * c + lane;                       // tile-relative position
                const uint32_t pabs = t0 + (uint32_t)i;
                const uint32_t x = in_s[kHist + i];
                s1 += x;                                            // Adler-32 partial sums (bytes past the end are zero)
                s2 += x * (n_tile - (uint32_t)i);
                // a match may start at p iff p >= 1 and p + 3 <= L - 2 (cur_search >= 0, di < isize - 3; deflate.py:975-977)
                const bool can = pabs >= 1u && pabs + 5u <= L;
                const uint32_t dmax = can ? (pabs < (uint32_t)kHist ? pabs : (uint32_t)kHist) : 0u;
                const uint32_t me = Wt[kHist + i];
                uint32_t d = 0;
                for (uint32_t dd = 1; __any_sync(HDLZ_FULL_MASK, d == 0u && dd <= dmax); ++dd)
                    if (d == 0u && dd <= dmax && Wt[kHist + i - (int)dd] == me) d = dd;
                uint32_t tk = literal_token(x);
                if (d) {
                    uint32_t m = 3;                                  // `di < isize - more`: p + m + 1 <= L - 2 (deflate.py:1043-1046)
                    while (m < (uint32_t)kMaxMatch && pabs + m + 3u <= L && in_s[kHist + i - (int)d + (int)m] == in_s[kHist + i + (int)m]) ++m;
                    tk = match_token(d, m);
                }
                if ((uint32_t)i >= n_tile) tk = 0;                  // positions past the end of the stream emit nothing
                Rw[i + (i >> 5)] = tk;
            }
            {
                const uint32_t S1 = __reduce_add_sync(HDLZ_FULL_MASK, s1);
                const uint32_t S2 = __reduce_add_sync(HDLZ_FULL_MASK, s2);
                adler_b = (adler_b + n_tile * adler_a + S2) % 65521u;
                adler_a = (adler_a + S1) % 65521u;
            }
            __syncwarp();

            // ---------------- P1: parse DP of the lane's segment, walked from the back
            unsigned long long H = 0;
            {
                const int rbase = 33 * lane;
#pragma unroll 8
                for (int j = kSeg - 1; j >= 0; --j) {
                    const uint32_t ls = Rw[rbase + j] >> 24;         // 4 * (length - 1)
                    const uint32_t look = (uint32_t)(H >> ls) & 15u;
                    const int ex = j + 1 + (int)(ls >> 2) - kSeg;
                    H = (H << 4) | (ex >= 0 ? (uint32_t)ex : look);
                }
            }

            // ---------------- P2: entry skip count of every segment
            uint32_t entry = 0;
            {
                unsigned long long *Hs = reinterpret_cast<unsigned long long *>(ws.stage);    // input bytes are dead
                __syncwarp();
                Hs[lane] = H;
                __syncwarp();
                uint32_t cur = carry;
#pragma unroll 8
                for (int s = 0; s < 32; ++s) {
                    if (lane == s) entry = cur;
                    cur = (uint32_t)(Hs[s] >> (4 * cur)) & 15u;
                }
                carry = cur;
                __syncwarp();
            }

            // ---------------- P3: this lane's tokens -> its private bitstream
            uint32_t nbits;
            {
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
                        acc |= (unsigned long long)(tke & 0x3FFFFu) << fill;
                        fill += (tke >> 18) & 31u;
                        r = start ? (tk >> 26) : r - 1;
                    }
                    if (fill >= 32) {                        // fill < 32 + 2 * 18 <= 64 between checks
                        priv[wcnt * 32 + lane] = (uint32_t)acc;
                        ++wcnt;
                        acc >>= 32;
                        fill -= 32;
                    }
                }
                if (fill) priv[wcnt * 32 + lane] = (uint32_t)acc;
                nbits = 32 * wcnt + fill;
            }
            uint32_t incl = nbits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(HDLZ_FULL_MASK, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t tile_bits = __shfl_sync(HDLZ_FULL_MASK, incl, 31);

            __syncwarp();
            for (int k = lane; k < kOutWords / 4; k += 32) reinterpret_cast<uint4 *>(outw)[k] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (lane == 0) outw[0] = pw;
            __syncwarp();

            // ---------------- merge: private streams -> the tile's stream at their bit offsets
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
                    const uint32_t val = __funnelshift_l(prev, cur, sh);
                    prev = cur;
                    if (i < nwo) {
                        if ((i == 0 && sh != 0) || (i == nwo - 1 && tail_partial)) atomicOr(&outw[w0 + i], val);
                        else outw[w0 + i] = val;
                    }
                }
            }
            __syncwarp();

            // ---------------- flush
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
                uint32_t trailer = 4;
                if (lane == 0) {
                    uint8_t *ob = reinterpret_cast<uint8_t *>(outw);
                    if (container == HDLZ_CONTAINER_ZLIB) {
                        ob[nbytes + 0] = (uint8_t)(adler_b >> 8);    // Adler-32 big-endian (deflate.py:788-814)
                        ob[nbytes + 1] = (uint8_t)(adler_b & 255);
                        ob[nbytes + 2] = (uint8_t)(adler_a >> 8);
                        ob[nbytes + 3] = (uint8_t)(adler_a & 255);
                    } else if (container == HDLZ_CONTAINER_GZIP) {
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
                for (uint32_t k = lane; k < nwords; k += 32) dst32[wbase + k] = outw[k];
                if (lane == 0) {
                    out_len[sid] = 4 * wbase + nbytes + trailer;
                    if (status) status[sid] = HDLZ_OK;
                }
            }
        }
        sid = n_warps + __shfl_sync(HDLZ_FULL_MASK, next_ticket, 0);
    }
}

}  // namespace

int launch_compress_wide(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                         uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                         uint32_t *d_status, uint64_t n, unsigned long long *queue, cudaStream_t s)
{
    if (!ctx->wide_attr_set) {
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress_wide<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        HDLZ_CUDA(cudaFuncSetAttribute(k_compress_wide<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        ctx->wide_attr_set = true;
    }
    uint64_t blocks = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint64_t resident = (uint64_t)ctx->sm_count * kCtasPerSm;
    if (blocks > resident) blocks = resident;
    if (ctx->max_match == 5)
        k_compress_wide<5><<<(unsigned)blocks, kWarpsPerCta * 32, kSmemBytes, s>>>(d_in, in_stride, d_in_len, uniform_len, d_out,
                                                                                    out_stride, d_out_len, d_status, n, queue,
                                                                                    ctx->container);
    else
        k_compress_wide<10><<<(unsigned)blocks, kWarpsPerCta * 32, kSmemBytes, s>>>(d_in, in_stride, d_in_len, uniform_len, d_out,
                                                                                     out_stride, d_out_len, d_status, n, queue,
                                                                                     ctx->container);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
