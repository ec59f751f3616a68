// hdlz_inflate.cu — zlib/deflate decompressor for sm_100a (stored, fixed and dynamic blocks,
// 32 KiB window), byte-exact with zlib.  Replaces the reference's decode states
// HEADER, BL, READBL/REPEAT, HF1..HF4/SPREAD, NEXT, INFLATE, D_NEXT, COPY
// (deflate.py:656-732, 1084-1659) — not a port of them:
//
//   one WARP per stream.  The bit reader, the Huffman state and the output cursor are
//   warp-uniform (every lane holds the same values), so symbol decode has no divergence and
//   table probes are shared-memory broadcasts; the lanes split up the wide work:
//     - input is fetched a 128-byte line at a time (one coalesced load, a word per lane,
//       the next line prefetched) and words are handed out with a shuffle — the analogue of
//       fill_buf's b1..b10 prefetch (deflate.py:423-515);
//     - canonical-Huffman tables (deflate.py HF1INIT..HF4, :1227-1380) are built cooperatively:
//       histogram by atomics, per-length ranks by match.any, replicated table fill per lane.
//       Unlike the reference there is no 32768-entry zero fill (HF1, :1204-1225): a 10-bit
//       primary table (2 KiB) plus a canonical slow path for codes longer than 10 bits;
//     - LZ copies (COPY, :1627-1656) are done 32 bytes per step, with the period trick for
//       overlapping copies (distance < length).
//   The fixed-tree table is built once per CTA and shared (the reference's `prev_method == 1`
//   shortcut, deflate.py:701-707, generalised).
//
// Output goes straight to HBM through L1; back-references read it back through L1 (same SM,
// ordered by __syncwarp).  Algorithmic HBM traffic per stream: C bytes read + L bytes written.

#include "hdlz_common.cuh"
#include "hdlz_frame.cuh"

namespace hdlz {
namespace {

constexpr int kWarps = 8;
constexpr int kLitBits = 10;
constexpr int kDistBits = 9;
constexpr int kClBits = 7;

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                         43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                        3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                          257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                          12289, 16385, 24577};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// One canonical Huffman code: primary lookup table + canonical arrays for the slow path.
struct Huff {
    uint16_t *tbl;        // 1 << tbits entries: (symbol << 4) | length, 0 = not in the primary table
    uint16_t *sorted;     // symbols ordered by (length, symbol)
    uint16_t *count;      // [16] codes per length
    int tbits;
};

struct __align__(16) WarpSmem {
    uint16_t lit_tbl[1 << kLitBits];
    uint16_t dist_tbl[1 << kDistBits];
    uint16_t cl_tbl[1 << kClBits];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t cl_sorted[20];
    uint16_t lit_count[16], dist_count[16], cl_count[16];
    uint16_t first[16], run[16], offs[16];   // scratch of build()
    uint8_t lens[320 + 16];
};

struct __align__(16) FixedSmem {
    uint16_t lit_tbl[1 << 9];
    uint16_t dist_tbl[1 << 5];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
};

// ---- bit reader: warp-uniform 64-bit window fed by shuffles from per-lane line registers ------
struct Reader {
    const uint8_t *base;   // stream start rounded down to 4 bytes
    uint32_t mis;          // stream start - base
    uint32_t end;          // mis + stream length (byte offset from base one past the stream)
    uint32_t wi;           // next 32-bit word (from base) to hand out
    uint32_t ca, cb;       // this lane's word of the current / next 32-word line
    uint64_t acc;
    uint32_t fill;
    int lane;

    __device__ __forceinline__ uint32_t load_line(uint32_t line) const
    {
        const uint32_t bo = (line * 32 + (uint32_t)lane) * 4;
        if (bo >= mis && bo + 4 <= end) return *reinterpret_cast<const uint32_t *>(base + bo);
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (bo + b >= mis && bo + b < end) w |= (uint32_t)base[bo + b] << (8 * b);
        return w;
    }
    __device__ __forceinline__ uint32_t next_word()
    {
        const uint32_t w = __shfl_sync(HDLZ_FULL_MASK, ca, wi & 31);
        ++wi;
        if ((wi & 31) == 0) {
            ca = cb;
            cb = load_line((wi >> 5) + 1);
        }
        return w;
    }
    // position the reader at byte `off` of the stream
    __device__ __forceinline__ void seek(uint32_t off)
    {
        const uint32_t a = mis + off;
        wi = a >> 2;
        ca = load_line(wi >> 5);
        cb = load_line((wi >> 5) + 1);
        const uint32_t sh = 8 * (a & 3);
        acc = (uint64_t)(next_word() >> sh);
        fill = 32 - sh;
    }
    // position the reader at bit `b` of the stream
    __device__ __forceinline__ void seek_bits(uint64_t b)
    {
        seek((uint32_t)(b >> 3));                // leaves fill >= 8
        if (b & 7u) drop((uint32_t)(b & 7u));
    }
    __device__ __forceinline__ void refill()     // call when fill < 32; afterwards fill >= 32
    {
        acc |= (uint64_t)next_word() << fill;
        fill += 32;
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)acc & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) { acc >>= n; fill -= n; }
    __device__ __forceinline__ uint32_t get(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    // bits of the stream consumed so far
    __device__ __forceinline__ int64_t bitpos() const { return (int64_t)wi * 32 - fill - 8 * (int64_t)mis; }
};

// Build a canonical Huffman decoder from `lens[0..nsym)`.  Returns 0, or 1 if the code is
// over-subscribed / illegally incomplete (zlib's inflate_table rules).  Warp-cooperative.
__device__ int build(const Huff &h, const uint8_t *lens, int nsym, bool allow_incomplete, WarpSmem &ws, int lane)
{
    const int tsize = 1 << h.tbits;
    for (int i = lane; i < tsize / 2; i += 32) reinterpret_cast<uint32_t *>(h.tbl)[i] = 0;
    if (lane < 16) { h.count[lane] = 0; ws.run[lane] = 0; }
    __syncwarp();
    for (int s = lane; s < nsym; s += 32) {
        // histogram; 16-bit counters packed two per word -> add into the right half
        const uint32_t l = lens[s];
        atomicAdd(reinterpret_cast<uint32_t *>(h.count) + (l >> 1), (l & 1) ? 0x10000u : 1u);
    }
    __syncwarp();
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        const int c = h.count[l];
        left = 2 * left - c;
        if (c) maxlen = l;
        if (left < 0) return 1;
    }
    if (maxlen == 0) return 0;               // no codes: every probe fails later (zlib: "invalid code" on use)
    if (left > 0 && !(allow_incomplete && maxlen == 1)) return 1;
    if (lane == 0) {
        uint32_t code = 0, off = 0;
        for (int l = 1; l <= 15; ++l) {
            ws.first[l] = (uint16_t)code;
            ws.offs[l] = (uint16_t)off;
            code = (code + h.count[l]) << 1;
            off += h.count[l];
        }
    }
    __syncwarp();
    for (int s0 = 0; s0 < nsym; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t l = s < nsym ? lens[s] : 0;
        const uint32_t peers = __match_any_sync(HDLZ_FULL_MASK, l);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        const uint32_t k = ws.run[l] + rank;           // index among the symbols of length l
        __syncwarp();
        if (rank == 0) ws.run[l] = (uint16_t)(ws.run[l] + __popc(peers));
        __syncwarp();
        if (l) {
            h.sorted[ws.offs[l] + k] = (uint16_t)s;
            if ((int)l <= h.tbits) {
                const uint32_t code = ws.first[l] + k;
                const uint32_t rev = __brev(code) >> (32 - l);
                const uint16_t ent = (uint16_t)((s << 4) | l);
                for (uint32_t idx = rev; idx < (uint32_t)tsize; idx += 1u << l) h.tbl[idx] = ent;
            }
        }
    }
    __syncwarp();
    return 0;
}

// Decode one symbol (warp-uniform).  Returns the symbol or -1 for an invalid code.
__device__ __forceinline__ int decode(Reader &r, const Huff &h)
{
    const uint32_t e = h.tbl[r.peek(h.tbits)];
    if (e & 15u) {
        r.drop(e & 15u);
        return (int)(e >> 4);
    }
    // slow path: canonical decode one bit at a time (codes longer than the primary table)
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
        code |= (int)((r.acc >> (l - 1)) & 1u);
        const int cnt = h.count[l];
        if (code - cnt < first) {
            r.drop(l);
            return h.sorted[index + (code - first)];
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// kStream: ONE stream whose input is still arriving (hdlz_dstream_*).  `received` bytes of it are in `in`; unless
// `final_input` says these are all, the launch stops where the next step could run out of input — before a block
// header it cannot see whole, or before a symbol when fewer than 64 bits are left — and records in *ctl what the
// reference's FSM keeps between clocks: the bit cursor (`di` / `dio`), the output cursor (`do`), where the
// current block's header started (its tables are rebuilt from there on the next launch) and BFINAL.  The next
// launch carries on from the record; the output so far stays in `out`, which is also the window.  This is the
// reference's decompressor waiting at `di >= isize - 4` until more input or IDLE (deflate.py:1529).  Compiled
// out of the batch kernel.
constexpr uint32_t kHeaderReserve = 320;     // bytes a dynamic block header can take (17 + 19 * 3 + 316 * 7 bits)
constexpr uint32_t kSymbolReserve = 64;      // bits: one length / distance pair with its extra bits is at most 48

template <bool kStream>
__global__ void __launch_bounds__(kWarps * 32)
k_inflate(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint64_t in_stride,
          const uint32_t *__restrict__ in_len, uint8_t *out, uint64_t out_stride, uint32_t out_cap,
          uint32_t *__restrict__ out_len, uint32_t *__restrict__ status, uint64_t n_streams, uint32_t flags,
          const uint32_t *__restrict__ work_list, const uint32_t *__restrict__ work_count, InflateCtl *ctl,
          uint32_t received, uint32_t final_input)
{
    // Work items: every stream (work_list == nullptr) or the streams the lane-per-stream kernel
    // handed over (dynamic blocks, unaligned buffers).  Persistent: warps stride over the items.
    const uint64_t n_items = work_list ? (uint64_t)*work_count : n_streams;
    if (n_items == 0) return;
    __shared__ WarpSmem s_warp[kWarps];
    __shared__ FixedSmem s_fixed;
    __shared__ uint32_t s_nib[16];        // CRC-32 nibble table (gzip trailer check)
    if (threadIdx.x < 16) s_nib[threadIdx.x] = crc32_nibble_entry(threadIdx.x);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem &ws = s_warp[warp];

    // fixed Huffman tables, once per CTA (STATIC, deflate.py:1064-1076)
    if (warp == 0) {
        for (int i = lane; i < 288; i += 32) ws.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        __syncwarp();
        Huff fl = {s_fixed.lit_tbl, s_fixed.lit_sorted, s_fixed.lit_count, 9};
        build(fl, ws.lens, 288, false, ws, lane);
        for (int i = lane; i < 32; i += 32) ws.lens[i] = 5;
        __syncwarp();
        Huff fd = {s_fixed.dist_tbl, s_fixed.dist_sorted, s_fixed.dist_count, 5};
        build(fd, ws.lens, 32, false, ws, lane);
    }
    __syncthreads();

    for (uint64_t item = (uint64_t)blockIdx.x * kWarps + warp; item < n_items; item += (uint64_t)gridDim.x * kWarps) {
    const uint64_t sid = work_list ? (uint64_t)work_list[item] : item;
    __syncwarp();

    const uint32_t n_in = kStream ? received : in_len[sid];
    const uint8_t *src = in + (in_off ? in_off[sid] : sid * in_stride);
    uint8_t *dst = out + sid * out_stride;

    uint32_t st = HDLZ_OK;
    uint32_t o = 0;
    // Literals are parked in registers — the byte for output position p sits in lane p % 32 — and
    // go to memory 32 at a time (one coalesced store) or right before something reads the output.
    uint32_t of = 0;                 // positions below `of` are in memory; of <= o <= of + 32
    uint32_t pend = 0;
    auto flush_literals = [&]() {
        const uint32_t p = of + (((uint32_t)lane - of) & 31u);
        if (p < o) dst[p] = (uint8_t)pend;
        of = o;
    };

    Reader r;
    r.lane = lane;
    r.mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    r.base = src - r.mis;
    r.end = r.mis + n_in;

    // container: zlib (the reference's, header skipped: di = 2, deflate.py:644), raw deflate or gzip
    const Frame frame = parse_frame(src, n_in, flags);
    st = frame.status;
    bool suspended = false;          // kStream: this launch stops short of the end of the stream
    if (kStream && !final_input && st == HDLZ_ST_TRUNCATED) {        // the container header has not arrived yet
        st = HDLZ_OK;
        suspended = true;
    }

    if (st == HDLZ_OK && !suspended) {
        const int64_t limit = 8 * (int64_t)n_in;
        const uint32_t wi_guard = (r.end + 8) / 4 + 2;   // words past the stream end: stop a runaway decode
        uint32_t final_blk = 0;
        bool in_block = false;       // kStream: resuming inside a Huffman block
        if (kStream && ctl->started) {
            o = ctl->o;
            of = o;
            in_block = !ctl->at_header;
            r.seek_bits(in_block ? ctl->hdr_bitpos : ctl->bitpos);
        } else {
            r.seek(frame.body);
        }
        auto suspend = [&](bool at_header, int64_t bits, int64_t hdr_bits) {
            suspended = true;
            if (lane == 0) {
                ctl->started = 1;
                ctl->at_header = at_header ? 1u : 0u;
                ctl->bitpos = (uint64_t)bits;
                ctl->hdr_bitpos = (uint64_t)hdr_bits;
                ctl->final_blk = final_blk;
            }
        };
        do {
            const int64_t hdr_bits = r.bitpos();
            if (kStream && !final_input && hdr_bits + 8 * (int64_t)kHeaderReserve > limit) {
                suspend(true, hdr_bits, hdr_bits);
                break;
            }
            if (r.fill < 32) r.refill();
            final_blk = r.get(1);
            const uint32_t type = r.get(2);
            if (type == 0) {
                // ---- stored block (deflate.py:709-717; COPY :1603-1616) ----
                const int64_t bp = r.bitpos();
                const uint32_t byte = (uint32_t)((bp + 7) >> 3);
                if ((uint64_t)byte + 4 > n_in) { st = HDLZ_ST_TRUNCATED; break; }
                const uint32_t len = src[byte] | ((uint32_t)src[byte + 1] << 8);
                const uint32_t nlen = src[byte + 2] | ((uint32_t)src[byte + 3] << 8);
                if ((len ^ 0xFFFFu) != nlen) { st = HDLZ_ST_BAD_STORED; break; }
                if (kStream && !final_input && (uint64_t)byte + 4 + len > n_in) {     // the block's bytes are still arriving
                    suspend(true, hdr_bits, hdr_bits);
                    break;
                }
                if ((uint64_t)byte + 4 + len > n_in) { st = HDLZ_ST_TRUNCATED; break; }
                if ((uint64_t)o + len > out_cap) { st = HDLZ_ST_OUT_OVERFLOW; break; }
                flush_literals();
                for (uint32_t k = lane; k < len; k += 32) dst[o + k] = src[byte + 4 + k];
                o += len;
                of = o;
                r.seek(byte + 4 + len);
                continue;
            }
            if (type == 3) { st = HDLZ_ST_BAD_BTYPE; break; }   // "Bad method" (deflate.py:718-721)

            Huff hl, hd;
            if (type == 1) {
                hl = Huff{s_fixed.lit_tbl, s_fixed.lit_sorted, s_fixed.lit_count, 9};
                hd = Huff{s_fixed.dist_tbl, s_fixed.dist_sorted, s_fixed.dist_count, 5};
            } else {
                // ---- dynamic block header (BL / READBL / REPEAT, deflate.py:1084-1202) ----
                if (r.fill < 32) r.refill();
                const uint32_t nlen = r.get(5) + 257, ndist = r.get(5) + 1, ncode = r.get(4) + 4;
                if (nlen > 286 || ndist > 30) { st = HDLZ_ST_BAD_CODE; break; }
                if (lane < 19) ws.lens[lane] = 0;
                __syncwarp();
                for (uint32_t i = 0; i < ncode; ++i) {
                    if (r.fill < 32) r.refill();
                    const uint32_t v = r.get(3);
                    if (lane == 0) ws.lens[c_cl_order[i]] = (uint8_t)v;
                }
                __syncwarp();
                Huff hc = {ws.cl_tbl, ws.cl_sorted, ws.cl_count, kClBits};
                // the code-length code must be complete (zlib: "invalid code lengths set")
                if (build(hc, ws.lens, 19, false, ws, lane)) { st = HDLZ_ST_BAD_CODE; break; }
                {
                    // zlib also rejects an empty code-length code
                    int any = 0;
                    for (int l = 1; l <= 7; ++l) any |= hc.count[l];
                    if (!any) { st = HDLZ_ST_BAD_CODE; break; }
                }
                __syncwarp();
                uint32_t idx = 0, prev = 0;
                const uint32_t total = nlen + ndist;
                // lengths are staged after the 19 code-length lengths are no longer needed
                while (idx < total) {
                    if (r.fill < 32) r.refill();
                    const int sym = decode(r, hc);
                    if (sym < 0) { st = HDLZ_ST_BAD_CODE; break; }
                    uint32_t rep, val;
                    if (sym < 16) { rep = 1; val = (uint32_t)sym; prev = val; }
                    else if (sym == 16) {
                        if (idx == 0) { st = HDLZ_ST_BAD_CODE; break; }
                        rep = 3 + r.get(2); val = prev;
                    } else if (sym == 17) { rep = 3 + r.get(3); val = 0; prev = 0; }
                    else { rep = 11 + r.get(7); val = 0; prev = 0; }
                    if (idx + rep > total) { st = HDLZ_ST_BAD_CODE; break; }
                    for (uint32_t k = lane; k < rep; k += 32) ws.lens[16 + idx + k] = (uint8_t)val;
                    idx += rep;
                }
                if (st != HDLZ_OK) break;
                __syncwarp();
                const uint8_t *ll = ws.lens + 16;
                if (ll[256] == 0) { st = HDLZ_ST_BAD_CODE; break; }        // no end-of-block code
                hl = Huff{ws.lit_tbl, ws.lit_sorted, ws.lit_count, kLitBits};
                hd = Huff{ws.dist_tbl, ws.dist_sorted, ws.dist_count, kDistBits};
                if (build(hl, ll, (int)nlen, true, ws, lane)) { st = HDLZ_ST_BAD_CODE; break; }
                if (build(hd, ll + nlen, (int)ndist, true, ws, lane)) { st = HDLZ_ST_BAD_CODE; break; }
            }

            if (kStream && in_block) {                 // tables rebuilt: back to where the previous launch stopped
                r.seek_bits(ctl->bitpos);
                in_block = false;
            }
            // ---- symbol loop (NEXT / INFLATE / D_NEXT / COPY, deflate.py:1402-1659) ----
            for (;;) {
                if (kStream && !final_input && r.bitpos() + (int64_t)kSymbolReserve > limit) {
                    flush_literals();
                    suspend(false, r.bitpos(), hdr_bits);
                    break;
                }
                if (r.fill < 32) {
                    r.refill();
                    if (r.wi > wi_guard) { st = HDLZ_ST_TRUNCATED; break; }
                }
                int sym = decode(r, hl);
                if (sym < 0) { st = HDLZ_ST_BAD_CODE; break; }
                if (sym < 256) {
                    if (o >= out_cap) { st = HDLZ_ST_OUT_OVERFLOW; break; }
                    if ((uint32_t)lane == (o & 31u)) pend = (uint32_t)sym;
                    ++o;
                    if (o - of == 32u) flush_literals();
                    continue;
                }
                if (sym == 256) break;
                sym -= 257;
                if (sym >= 29) { st = HDLZ_ST_BAD_CODE; break; }           // "invalid token" (deflate.py:1559-1560)
                const uint32_t len = c_len_base[sym] + r.get(c_len_extra[sym]);
                if (r.fill < 32) r.refill();
                const int dsym = decode(r, hd);
                if (dsym < 0 || dsym >= 30) { st = HDLZ_ST_BAD_CODE; break; }
                const uint32_t dbits = dsym < 2 ? 0u : (uint32_t)(dsym >> 1) - 1u;
                const uint32_t dist = c_dist_base[dsym] + r.get(dbits);
                if (dist > o) { st = HDLZ_ST_DIST_TOO_FAR; break; }        // "distance too big" (deflate.py:1506-1508)
                if ((uint64_t)o + len > out_cap) { st = HDLZ_ST_OUT_OVERFLOW; break; }
                flush_literals();
                __syncwarp();
                if (dist >= len) {
                    for (uint32_t k = lane; k < len; k += 32) dst[o + k] = dst[o - dist + k];
                } else {
                    for (uint32_t k = lane; k < len; k += 32) dst[o + k] = dst[o - dist + (k % dist)];
                }
                __syncwarp();
                o += len;
                of = o;
            }
            if (st != HDLZ_OK || suspended) break;
            if (r.bitpos() > limit) { st = HDLZ_ST_TRUNCATED; break; }
        } while (!final_blk);

        flush_literals();
        if (st == HDLZ_OK && !suspended) {
            // the trailer (zlib: Adler-32, four bytes) must exist after the next byte boundary ("NO EOF!",
            // deflate.py:1535-1539)
            const int64_t bp = r.bitpos();
            const uint32_t tp = (uint32_t)((bp + 7) >> 3);
            if (bp > limit || (uint64_t)tp + frame.trailer > n_in) {
                st = HDLZ_ST_TRUNCATED;
            } else if ((flags & HDLZ_F_VERIFY_ADLER) && (flags & HDLZ_F_GZIP)) {
                // gzip: CRC-32 and ISIZE (RFC 1952); one lane walks the output
                __syncwarp();
                uint32_t bad = 0;
                if (lane == 0) bad = crc32_bytes(dst, o, s_nib) != load_le32(src + tp) || o != load_le32(src + tp + 4);
                if (__shfl_sync(HDLZ_FULL_MASK, bad, 0)) st = HDLZ_ST_BAD_CRC;
            } else if ((flags & HDLZ_F_VERIFY_ADLER) && !(flags & HDLZ_F_RAW)) {
                __syncwarp();
                uint64_t s1 = 0, s2 = 0;
                for (uint32_t i = lane; i < o; i += 32) {
                    const uint32_t v = dst[i];
                    s1 += v;
                    s2 += (uint64_t)v * (o - i);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    s1 += __shfl_xor_sync(HDLZ_FULL_MASK, s1, d);
                    s2 += __shfl_xor_sync(HDLZ_FULL_MASK, s2, d);
                }
                const uint32_t a = (uint32_t)((1 + s1) % 65521u);
                const uint32_t b = (uint32_t)((o + s2) % 65521u);
                const uint32_t want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) |
                                      ((uint32_t)src[tp + 2] << 8) | src[tp + 3];
                if (((b << 16) | a) != want) st = HDLZ_ST_BAD_ADLER;
            }
        }
    }

    if (kStream) {
        __syncwarp();
        if (lane == 0) {
            ctl->o = o;                          // bytes in `out` so far (all written)
            ctl->done = suspended ? 0u : 1u;
            ctl->status = st;
        }
    } else if (lane == 0) {
        out_len[sid] = st == HDLZ_OK ? o : 0;
        if (status) status[sid] = st;
    }
    }  // work items
}

}  // namespace

int launch_inflate_general(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                           const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                           uint32_t *d_out_len, uint32_t *d_status, uint64_t n, uint32_t flags,
                           const uint32_t *work_list, const uint32_t *work_count, cudaStream_t s)
{
    if (n == 0) return HDLZ_SUCCESS;
    uint64_t blocks = (n + kWarps - 1) / kWarps;
    const uint64_t resident = (uint64_t)ctx->sm_count * 5;       // 38 KiB of shared memory per CTA
    if (blocks > resident) blocks = resident;
    k_inflate<false><<<(unsigned)blocks, kWarps * 32, 0, s>>>(d_in, d_in_off, in_stride, d_in_len, d_out, out_stride, out_cap,
                                                               d_out_len, d_status, n, flags, work_list, work_count, nullptr, 0, 0);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

// One launch of the stream kernel (hdlz_dstream_feed / hdlz_dstream_finish, hdlz_api.cu): a single warp works.
int launch_inflate_stream(hdlz_ctx *ctx, const uint8_t *d_in, uint32_t received, bool final_input, uint8_t *d_out,
                          uint32_t out_cap, uint32_t flags, InflateCtl *d_ctl, cudaStream_t s)
{
    k_inflate<true><<<1, kWarps * 32, 0, s>>>(d_in, nullptr, 0, nullptr, d_out, 0, out_cap, nullptr, nullptr, 1, flags, nullptr,
                                              nullptr, d_ctl, received, final_input ? 1u : 0u);
    ctx->launches++;
    HDLZ_CUDA(cudaGetLastError());
    return HDLZ_SUCCESS;
}

}  // namespace hdlz
