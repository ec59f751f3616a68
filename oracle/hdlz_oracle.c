/*
 * hdlz_oracle.c — TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
 *
 * A scalar CPU restatement of the reference's hot path (tomtor/HDL-deflate,
 * deflate.py), written from the behaviour of the reference FSM; every function
 * cites the reference lines it follows.  Nothing under hdl-deflate_b200/ may
 * link, load or call this file: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or as
 * the timed CPU arm.
 *
 * Parity status: PINNED.  The compress restatement is checked byte-for-byte
 * against the unmodified reference engine executed under the MyHDL-compat
 * layer (oracle/ref_sim.py) on the golden vectors of tests/golden/ (generated
 * by oracle/make_golden.py) and on seeded fuzz inputs (tests/test_oracle.py).
 * The inflate restatement is checked against zlib, which is the reference's own
 * decompress check (test_deflate.py:194).
 *
 * Build: see oracle/Makefile (gcc -O3 -shared -fPIC ... -lz -lpthread).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <zlib.h>

/* status codes: keep in sync with include/hdlz.h */
enum {
    ST_OK = 0, ST_SHORT_INPUT = 1, ST_BAD_BTYPE = 2, ST_BAD_CODE = 3, ST_DIST_TOO_FAR = 4,
    ST_TRUNCATED = 5, ST_OUT_OVERFLOW = 6, ST_BAD_STORED = 7, ST_BAD_HEADER = 8, ST_BAD_ADLER = 9
};
#define F_VERIFY_HEADER 1u
#define F_VERIFY_ADLER  2u

/* ------------------------------------------------------------------------ */
/* tables (RFC 1951 facts; deflate.py:100-110 holds the same values)         */
/* ------------------------------------------------------------------------ */
static const uint16_t kCopyLength[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                          43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kExtraLengthBits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                             3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kCopyDistance[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                            257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                            12289, 16385, 24577};
static const uint8_t kCodeLengthOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static unsigned rev_bits(unsigned v, unsigned n)   /* deflate.py:569-584 */
{
    unsigned r = 0;
    for (unsigned i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

/* Fixed-Huffman code of literal/length symbol s, bit-reversed for LSB-first
 * emission, and its length.  Equals out_codes[s] / codeLength[s] of the
 * reference (deflate.py:112-149, 1064-1076); generated from RFC 1951 3.2.6. */
static void fixed_code(unsigned s, unsigned *code, unsigned *len)
{
    if (s < 144)      { *len = 8; *code = rev_bits(0x30 + s, 8); }
    else if (s < 256) { *len = 9; *code = rev_bits(0x190 + (s - 144), 9); }
    else if (s < 280) { *len = 7; *code = rev_bits(s - 256, 7); }
    else              { *len = 8; *code = rev_bits(0xC0 + (s - 280), 8); }
}

/* ------------------------------------------------------------------------ */
/* LSB-first bit writer  (put/do_flush, deflate.py:535-567)                  */
/* ------------------------------------------------------------------------ */
typedef struct { uint8_t *out; uint32_t cap, pos; uint32_t acc; unsigned fill; int ovf; } bitw_t;

static void bw_put(bitw_t *w, unsigned v, unsigned n)
{
    w->acc |= v << w->fill;
    w->fill += n;
    while (w->fill >= 8) {
        if (w->pos < w->cap) w->out[w->pos] = (uint8_t)w->acc; else w->ovf = 1;
        w->pos++;
        w->acc >>= 8;
        w->fill -= 8;
    }
}

/* ------------------------------------------------------------------------ */
/* compress: FAST + MATCH10, CWINDOW=32, one fixed block per stream          */
/* ------------------------------------------------------------------------ */

/* Match rule of SEARCH + SEARCHF (deflate.py:966-994, 899-964) with the 32
 * matcher3 comparators (deflate.py:407-418): nearest distance whose 3 bytes
 * match wins, then the length grows to at most `maxlen` (10 with MATCH10, 5
 * without) while p + m + 1 <= L - 2  (the `di < isize - k` guards).          */
static int find_match(const uint8_t *x, uint32_t L, uint32_t p, unsigned cwindow, unsigned maxlen,
                      unsigned *dist, unsigned *len)
{
    if (p < 1 || p + 5 > L) return 0;          /* cur_search >= 0 and di < isize - 3 */
    unsigned dmax = p < cwindow ? p : cwindow;
    for (unsigned d = 1; d <= dmax; d++) {     /* lowest si first => nearest (:982-988) */
        if (x[p - d] == x[p] && x[p - d + 1] == x[p + 1] && x[p - d + 2] == x[p + 2]) {
            unsigned m = 3;
            while (m < maxlen && p + m + 3 <= L && x[p - d + m] == x[p + m]) m++;
            *dist = d; *len = m;
            return 1;
        }
    }
    return 0;
}

/* Emission of one stream: header 78 9C, BFINAL=1/BTYPE=01 (CSTATIC, :746-762),
 * tokens (SEARCH literal :1004-1016, DISTANCE :836-882), EOB + pad + Adler-32
 * big-endian (:771-814).  Returns the status; *out_len is the stream length
 * (== o_oprogress when o_done rises).                                         */
int hdlz_oracle_compress_ex(const uint8_t *x, uint32_t L, uint8_t *out, uint32_t cap, uint32_t *out_len,
                            unsigned cwindow, unsigned maxlen)
{
    *out_len = 0;
    if (L < 5) return ST_SHORT_INPUT;          /* engine idles while isize < 4 (:429-432) */
    bitw_t w = {out, cap, 0, 0, 0, 0};
    bw_put(&w, 0x78, 8);
    bw_put(&w, 0x9C, 8);
    bw_put(&w, 3, 3);
    uint32_t a = 1, b = 0;
    uint32_t p = 0;
    while (p < L) {
        unsigned d, m, code, len;
        if (find_match(x, L, p, cwindow, maxlen, &d, &m)) {
            fixed_code(254 + m, &code, &len);              /* lencode = mlength + 254 (:845) */
            bw_put(&w, code, len);
            unsigned c = 0;
            while (kCopyDistance[c + 1] <= d) c++;          /* :858-860 */
            unsigned eb = c < 2 ? 0 : (c >> 1) - 1;         /* ExtraDistanceBits[c // 2] (:864) */
            unsigned oc = rev_bits(c, 5) | ((d - kCopyDistance[c]) << 5);   /* :870 */
            if (5 + eb <= 9) bw_put(&w, oc, 5 + eb);
            else { bw_put(&w, oc & 0xFF, 8); bw_put(&w, oc >> 8, eb - 3); }  /* outcarry (:875-880) */
        } else {
            m = 1;
            fixed_code(x[p], &code, &len);
            bw_put(&w, code, len);
        }
        for (unsigned k = 0; k < m; k++) {                  /* CSTATIC :826-831, CHECKSUM :884-897 */
            a = (a + x[p + k]) % 65521u;
            b = (b + a) % 65521u;
        }
        p += m;
    }
    bw_put(&w, 0, 7);                                       /* EOB symbol 256 (:772-779) */
    if (w.fill) bw_put(&w, 0, 8 - w.fill);                  /* flush partial byte (:784-787) */
    bw_put(&w, b >> 8, 8); bw_put(&w, b & 255, 8);          /* :788-814 */
    bw_put(&w, a >> 8, 8); bw_put(&w, a & 255, 8);
    *out_len = w.pos;
    return w.ovf ? ST_OUT_OVERFLOW : ST_OK;
}

int hdlz_oracle_compress(const uint8_t *x, uint32_t L, uint8_t *out, uint32_t cap, uint32_t *out_len)
{
    return hdlz_oracle_compress_ex(x, L, out, cap, out_len, 32, 10);
}

/* ------------------------------------------------------------------------ */
/* Tuned CPU arm of the same contract (the "fair" baseline of bench.py):     */
/* identical bytes to hdlz_oracle_compress (tests/test_oracle.py checks it   */
/* on the golden vectors and on fuzz), but written for speed — SSE2 byte     */
/* compares stand in for the 32 matcher3 comparators (deflate.py:407-418),   */
/* a 64-bit bit buffer for put/do_flush (:535-567), zlib's adler32 for the   */
/* CSTATIC / CHECKSUM sums (:826-831, :884-897).  CWINDOW = 32 only.         */
/* ------------------------------------------------------------------------ */
#include <emmintrin.h>

/* bit i <=> x[p - 32 + i] == x[p]  (distance 32 - i); needs p >= 32 */
static inline uint32_t eq_mask32(const uint8_t *x, uint32_t p)
{
    const __m128i v = _mm_set1_epi8((char)x[p]);
    const __m128i a = _mm_loadu_si128((const __m128i *)(x + p - 32));
    const __m128i b = _mm_loadu_si128((const __m128i *)(x + p - 16));
    return (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(a, v)) |
           ((uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(b, v)) << 16);
}

typedef struct { uint8_t *out; uint32_t cap, pos; uint64_t acc; unsigned fill; int ovf; } bitw64_t;

static inline void bw64_put(bitw64_t *w, uint32_t v, unsigned n)     /* n <= 24, fill < 32 on entry */
{
    w->acc |= (uint64_t)v << w->fill;
    w->fill += n;
    if (w->fill >= 32) {
        if (w->pos + 4 <= w->cap) memcpy(w->out + w->pos, &w->acc, 4); else w->ovf = 1;
        w->pos += 4;
        w->acc >>= 32;
        w->fill -= 32;
    }
}

int hdlz_oracle_compress_fast(const uint8_t *x, uint32_t L, uint8_t *out, uint32_t cap, uint32_t *out_len,
                              unsigned maxlen)
{
    static uint32_t lit_tok[256], len_tok[11], dist_tok[33];     /* code | nbits << 16 */
    static volatile int ready = 0;
    if (!ready) {
        for (unsigned s = 0; s < 256; s++) { unsigned c, n; fixed_code(s, &c, &n); lit_tok[s] = c | (n << 16); }
        for (unsigned m = 3; m <= 10; m++) { unsigned c, n; fixed_code(254 + m, &c, &n); len_tok[m] = c | (n << 16); }
        for (unsigned d = 1; d <= 32; d++) {
            unsigned c = 0;
            while (kCopyDistance[c + 1] <= d) c++;
            unsigned eb = c < 2 ? 0 : (c >> 1) - 1;
            dist_tok[d] = (rev_bits(c, 5) | ((d - kCopyDistance[c]) << 5)) | ((5 + eb) << 16);
        }
        __sync_synchronize();
        ready = 1;
    }
    *out_len = 0;
    if (L < 5) return ST_SHORT_INPUT;
    bitw64_t w = {out, cap, 0, 0, 0, 0};
    bw64_put(&w, 0x78 | (0x9C << 8) | (3u << 16), 19);
    uint32_t p = 0;
    while (p < L) {
        unsigned d = 0, m = 1;
        if (p >= 1 && p + 5 <= L) {
            if (p >= 32) {
                uint32_t m3 = eq_mask32(x, p) & eq_mask32(x, p + 1) & eq_mask32(x, p + 2);
                if (m3) d = (unsigned)__builtin_clz(m3) + 1;       /* highest bit = nearest distance */
            } else {
                for (unsigned dd = 1; dd <= p; dd++)
                    if (x[p - dd] == x[p] && x[p - dd + 1] == x[p + 1] && x[p - dd + 2] == x[p + 2]) { d = dd; break; }
            }
            if (d) {
                m = 3;
                while (m < maxlen && p + m + 3 <= L && x[p - d + m] == x[p + m]) m++;
            }
        }
        if (d) {
            bw64_put(&w, len_tok[m] & 0xFFFF, len_tok[m] >> 16);
            bw64_put(&w, dist_tok[d] & 0xFFFF, dist_tok[d] >> 16);
        } else {
            bw64_put(&w, lit_tok[x[p]] & 0xFFFF, lit_tok[x[p]] >> 16);
        }
        p += m;
    }
    bw64_put(&w, 0, 7);
    if (w.fill & 7) bw64_put(&w, 0, 8 - (w.fill & 7));
    const uint32_t ad = (uint32_t)adler32(1L, x, L);
    while (w.fill) {                                              /* whole bytes left in the buffer */
        if (w.pos < w.cap) w.out[w.pos] = (uint8_t)w.acc; else w.ovf = 1;
        w.pos++; w.acc >>= 8; w.fill -= 8;
    }
    for (int k = 3; k >= 0; k--) {
        if (w.pos < w.cap) w.out[w.pos] = (uint8_t)(ad >> (8 * k)); else w.ovf = 1;
        w.pos++;
    }
    *out_len = w.pos;
    return w.ovf ? ST_OUT_OVERFLOW : ST_OK;
}

/* Token trace for debugging / structural tests: fills tok[i] = p | len<<24 | dist<<16. */
uint32_t hdlz_oracle_parse(const uint8_t *x, uint32_t L, uint32_t *tok, uint32_t cap)
{
    uint32_t n = 0, p = 0;
    while (p < L) {
        unsigned d = 0, m = 1;
        if (!find_match(x, L, p, 32, 10, &d, &m)) { d = 0; m = 1; }
        if (n < cap) tok[n] = p | (m << 24) | (d << 16);
        n++;
        p += m;
    }
    return n;
}

/* The same greedy parse (SEARCH / SEARCHF, deflate.py:899-1016) as (length, distance) pairs, for streams of any
 * length and either MATCH10 setting: tok[2i] = length (1 = literal), tok[2i+1] = distance.  Used by the tree-mode
 * restatement (oracle/tree_oracle.py). */
uint32_t hdlz_oracle_parse_ex(const uint8_t *x, uint32_t L, uint32_t *tok, uint32_t cap, unsigned cwindow, unsigned maxlen)
{
    uint32_t n = 0, p = 0;
    while (p < L) {
        unsigned d = 0, m = 1;
        if (!find_match(x, L, p, cwindow, maxlen, &d, &m)) { d = 0; m = 1; }
        if (n < cap) { tok[2 * n] = m; tok[2 * n + 1] = d; }
        n++;
        p += m;
    }
    return n;
}

/* ------------------------------------------------------------------------ */
/* inflate restatement (zlib-wrapped RFC 1951; HEADER..COPY, :656-732,       */
/* :1084-1659).  Parity target is zlib (test_deflate.py:194).                */
/* ------------------------------------------------------------------------ */
typedef struct { const uint8_t *in; uint32_t len, pos; uint64_t acc; unsigned fill; } bitr_t;

static int br_need(bitr_t *r, unsigned n)
{
    while (r->fill < n) {
        if (r->pos >= r->len) return 0;
        r->acc |= (uint64_t)r->in[r->pos++] << r->fill;
        r->fill += 8;
    }
    return 1;
}
static int br_get(bitr_t *r, unsigned n, unsigned *v)
{
    if (n == 0) { *v = 0; return 1; }
    if (!br_need(r, n)) return 0;
    *v = (unsigned)(r->acc & ((1u << n) - 1));
    r->acc >>= n; r->fill -= n;
    return 1;
}

typedef struct { uint16_t count[16]; uint16_t sym[320]; } huff_t;

/* canonical Huffman from code lengths (HF1INIT..HF4, :1227-1380).
 * returns 0 complete, <0 over-subscribed, >0 incomplete                     */
static int huff_build(huff_t *h, const uint8_t *lens, int n)
{
    uint16_t offs[16];
    memset(h->count, 0, sizeof h->count);
    for (int i = 0; i < n; i++) h->count[lens[i]]++;
    if (h->count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h->count[l];
    for (int i = 0; i < n; i++) if (lens[i]) h->sym[offs[lens[i]]++] = (uint16_t)i;
    return left;
}

static int huff_decode(bitr_t *r, const huff_t *h, unsigned *sym)   /* NEXT (:1402-1445) */
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        unsigned bit;
        if (!br_get(r, 1, &bit)) return ST_TRUNCATED;
        code |= (int)bit;
        int count = h->count[l];
        if (code - count < first) { *sym = h->sym[index + (code - first)]; return ST_OK; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return ST_BAD_CODE;
}

int hdlz_oracle_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t cap, uint32_t *out_len,
                        uint32_t flags)
{
    bitr_t r = {in, in_len, 0, 0, 0};
    uint32_t o = 0;
    *out_len = 0;
    if (in_len < 2) return ST_TRUNCATED;
    if (flags & F_VERIFY_HEADER) {
        unsigned cmf = in[0], flg = in[1];
        if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 || (flg & 0x20)) return ST_BAD_HEADER;
    }
    r.pos = 2;                                              /* di = 2 (:644) */
    unsigned final;
    do {
        unsigned type;
        if (!br_get(&r, 1, &final) || !br_get(&r, 2, &type)) return ST_TRUNCATED;
        if (type == 0) {                                    /* stored (:709-717, COPY :1603-1616) */
            r.acc = 0; r.fill = 0;                          /* to byte boundary: whole bytes only are buffered */
            /* re-sync byte position: bytes already pulled into acc beyond the boundary */
            /* (acc is refilled byte-wise, so pos is exact after dropping fill)         */
            if (r.pos + 4 > r.len) return ST_TRUNCATED;
            unsigned len = in[r.pos] | (in[r.pos + 1] << 8);
            unsigned nlen = in[r.pos + 2] | (in[r.pos + 3] << 8);
            r.pos += 4;
            if ((len ^ 0xFFFF) != nlen) return ST_BAD_STORED;
            if (r.pos + len > r.len) return ST_TRUNCATED;
            if (o + len > cap) return ST_OUT_OVERFLOW;
            memcpy(out + o, in + r.pos, len);
            o += len; r.pos += len;
            continue;
        }
        if (type == 3) return ST_BAD_BTYPE;                 /* "Bad method" (:718-721) */
        huff_t hl, hd;
        uint8_t lens[320];
        if (type == 1) {                                    /* STATIC (:1064-1076) */
            int i = 0;
            for (; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            huff_build(&hl, lens, 288);
            for (i = 0; i < 30; i++) lens[i] = 5;
            huff_build(&hd, lens, 30);
        } else {                                            /* BL / READBL / REPEAT (:1084-1202) */
            unsigned nlen, ndist, ncode, v;
            if (!br_get(&r, 5, &nlen) || !br_get(&r, 5, &ndist) || !br_get(&r, 4, &ncode)) return ST_TRUNCATED;
            nlen += 257; ndist += 1; ncode += 4;
            if (nlen > 286 || ndist > 30) return ST_BAD_CODE;
            memset(lens, 0, 19);
            for (unsigned i = 0; i < ncode; i++) {
                if (!br_get(&r, 3, &v)) return ST_TRUNCATED;
                lens[kCodeLengthOrder[i]] = (uint8_t)v;
            }
            huff_t hc;
            if (huff_build(&hc, lens, 19) != 0) return ST_BAD_CODE;
            unsigned idx = 0;
            while (idx < nlen + ndist) {
                unsigned sym, rep;
                int st = huff_decode(&r, &hc, &sym);
                if (st) return st;
                if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
                unsigned prev = 0;
                if (sym == 16) {
                    if (idx == 0) return ST_BAD_CODE;
                    prev = lens[idx - 1];
                    if (!br_get(&r, 2, &rep)) return ST_TRUNCATED;
                    rep += 3;
                } else if (sym == 17) {
                    if (!br_get(&r, 3, &rep)) return ST_TRUNCATED;
                    rep += 3;
                } else {
                    if (!br_get(&r, 7, &rep)) return ST_TRUNCATED;
                    rep += 11;
                }
                if (idx + rep > nlen + ndist) return ST_BAD_CODE;
                while (rep--) lens[idx++] = (uint8_t)prev;
            }
            if (lens[256] == 0) return ST_BAD_CODE;
            int e = huff_build(&hl, lens, (int)nlen);
            if (e && (e < 0 || nlen != (unsigned)(hl.count[0] + hl.count[1]))) return ST_BAD_CODE;
            e = huff_build(&hd, lens + nlen, (int)ndist);
            if (e && (e < 0 || ndist != (unsigned)(hd.count[0] + hd.count[1]))) return ST_BAD_CODE;
        }
        for (;;) {                                          /* NEXT / INFLATE / D_NEXT / COPY */
            unsigned sym, eb, dsym;
            int st = huff_decode(&r, &hl, &sym);
            if (st) return st;
            if (sym < 256) {
                if (o >= cap) return ST_OUT_OVERFLOW;
                out[o++] = (uint8_t)sym;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) return ST_BAD_CODE;              /* "invalid token" (:1559-1560) */
            if (!br_get(&r, kExtraLengthBits[sym], &eb)) return ST_TRUNCATED;
            unsigned len = kCopyLength[sym] + eb;
            st = huff_decode(&r, &hd, &dsym);
            if (st) return st;
            if (dsym >= 30) return ST_BAD_CODE;
            unsigned dbits = dsym < 2 ? 0 : (dsym >> 1) - 1;
            if (!br_get(&r, dbits, &eb)) return ST_TRUNCATED;
            unsigned dist = kCopyDistance[dsym] + eb;
            if (dist > o) return ST_DIST_TOO_FAR;           /* "distance too big" (:1506-1508) */
            if (o + len > cap) return ST_OUT_OVERFLOW;
            for (unsigned k = 0; k < len; k++, o++) out[o] = out[o - dist];
        }
    } while (!final);
    *out_len = o;
    if (flags & F_VERIFY_ADLER) {
        /* trailer sits at the next byte boundary; bytes buffered in acc are whole bytes */
        uint32_t tp = r.pos - r.fill / 8;
        if (tp + 4 > in_len) return ST_TRUNCATED;
        uint32_t want = ((uint32_t)in[tp] << 24) | (in[tp + 1] << 16) | (in[tp + 2] << 8) | in[tp + 3];
        uint32_t a = 1, b = 0;
        for (uint32_t i = 0; i < o; i++) { a = (a + out[i]) % 65521u; b = (b + a) % 65521u; }
        if (((b << 16) | a) != want) return ST_BAD_ADLER;
    }
    return ST_OK;
}

/* ------------------------------------------------------------------------ */
/* multi-threaded batch drivers (CPU baselines of bench.py)                  */
/* ------------------------------------------------------------------------ */
typedef struct {
    int kind;                 /* 0 port-compress, 1 zlib deflate, 2 zlib inflate, 3 port-inflate, 4 tuned port-compress */
    const uint8_t *in; const uint64_t *in_off; const uint32_t *in_len;
    uint8_t *out; const uint64_t *out_off; uint32_t out_cap; uint32_t *out_len; uint32_t *status;
    uint64_t lo, hi; int level, strategy;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    z_stream zs;
    int zinit = 0;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        const uint8_t *src = j->in + j->in_off[i];
        uint8_t *dst = j->out + j->out_off[i];
        uint32_t n = 0;
        int st = 0;
        if (j->kind == 0) {
            st = hdlz_oracle_compress(src, j->in_len[i], dst, j->out_cap, &n);
        } else if (j->kind == 4) {
            st = hdlz_oracle_compress_fast(src, j->in_len[i], dst, j->out_cap, &n, 10);
        } else if (j->kind == 3) {
            st = hdlz_oracle_inflate(src, j->in_len[i], dst, j->out_cap, &n, 0);
        } else if (j->kind == 1) {
            if (!zinit) {
                memset(&zs, 0, sizeof zs);
                deflateInit2(&zs, j->level, Z_DEFLATED, 15, 8, j->strategy);
                zinit = 1;
            } else deflateReset(&zs);
            zs.next_in = (Bytef *)src; zs.avail_in = j->in_len[i];
            zs.next_out = dst; zs.avail_out = j->out_cap;
            st = deflate(&zs, Z_FINISH) == Z_STREAM_END ? 0 : ST_OUT_OVERFLOW;
            n = (uint32_t)zs.total_out;
        } else {
            if (!zinit) { memset(&zs, 0, sizeof zs); inflateInit2(&zs, 15); zinit = 1; }
            else inflateReset(&zs);
            zs.next_in = (Bytef *)src; zs.avail_in = j->in_len[i];
            zs.next_out = dst; zs.avail_out = j->out_cap;
            st = inflate(&zs, Z_FINISH) == Z_STREAM_END ? 0 : ST_BAD_CODE;
            n = (uint32_t)zs.total_out;
        }
        j->out_len[i] = n;
        if (j->status) j->status[i] = (uint32_t)st;
    }
    if (zinit) { if (j->kind == 1) deflateEnd(&zs); else inflateEnd(&zs); }
    return NULL;
}

/* Generic batch: item i reads in[in_off[i] .. +in_len[i]) and writes out[out_off[i] .. +out_cap). */
int hdlz_oracle_batch(int kind, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                      uint8_t *out, const uint64_t *out_off, uint32_t out_cap, uint32_t *out_len,
                      uint32_t *status, uint64_t n, int nthreads, int level, int strategy)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        job_t j = {kind, in, in_off, in_len, out, out_off, out_cap, out_len, status,
                   n * t / nthreads, n * (t + 1) / nthreads, level, strategy};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
