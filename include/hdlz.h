/*
 * hdlz.h — C ABI of the B200-native deflate engine (libhdlz.so).
 *
 * This is the drop-in boundary for the hot path of tomtor/HDL-deflate: what a
 * host would bind by FFI (ctypes / cgo / JNI) instead of clocking the
 * reference's `deflate()` block.  The reference has no C interface — its
 * interface is the ten-signal port of deflate.py:220-221 driven one command
 * per clock (deflate.py:18, 599-605, 616-654) — so each entry point below names
 * the reference behaviour it replaces.  The Python host model that keeps the
 * reference's port protocol on top of these calls is
 * hdl-deflate_b200/dropin/deflate.py; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).
 *   - functions return 0 (HDLZ_SUCCESS) or a negative hdlz_error;
 *     hdlz_last_error() gives the text (thread-local).
 *   - per-stream outcomes are reported in `status[]` (hdlz_status), the C
 *     counterpart of the reference's `raise Error(...)` sites.
 *   - *_batch functions take DEVICE pointers and are asynchronous on `stream`;
 *     *_host / *_stream functions take HOST pointers, copy in, run the same
 *     kernels, copy out and synchronise before returning.
 *   - alignment: d_in, d_out 16-byte aligned; in_stride, out_stride multiples
 *     of 16 (cudaMalloc / torch allocations satisfy this).
 *   - the library never falls back to a CPU implementation: without a CUDA
 *     device every compute entry point fails with HDLZ_ERR_NODEVICE.
 */
#ifndef HDLZ_H
#define HDLZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDLZ_VERSION 0x000100 /* 0.1.0 */

/* Engine constants the reference exports as module globals (deflate.py:56-89). */
#define HDLZ_CWINDOW 32      /* search window, FAST (deflate.py:56-57)               */
#define HDLZ_CWINDOW_SLOW 256 /* search window with FAST = False (deflate.py:58-59)  */
#define HDLZ_MAX_MATCH 10    /* MATCH10 (deflate.py:34-35, 913-952)                  */
#define HDLZ_MAX_MATCH_SHORT 5 /* MATCH10 = False: SEARCHF stops at 5 (deflate.py:913-924) */
#define HDLZ_MIN_INPUT 5     /* engine idles while isize < 4 (deflate.py:429-432)    */
#define HDLZ_OBSIZE 32768    /* decompress window, "ALL valid streams" (README:20-21) */
#define HDLZ_LMAX 24         /* width of progress / address counters (deflate.py:73-76) */
#define HDLZ_LONG_STREAM 65536 /* hdlz_compress_stream: from this length on the stream is spread over the whole GPU */

typedef enum hdlz_error {
    HDLZ_SUCCESS = 0,
    HDLZ_ERR_INVALID = -1,  /* bad argument (null pointer, misaligned, stride too small) */
    HDLZ_ERR_CUDA = -2,     /* a CUDA runtime call failed (text in hdlz_last_error)      */
    HDLZ_ERR_NOMEM = -3,    /* device or pinned allocation failed                        */
    HDLZ_ERR_NODEVICE = -4  /* no CUDA device / driver                                   */
} hdlz_error;

/* Per-stream status word.  The text in quotes is the message the reference
 * raises at the cited line; dropin/deflate.py re-raises it as myhdl.Error.      */
typedef enum hdlz_status {
    HDLZ_OK = 0,
    HDLZ_ST_SHORT_INPUT = 1,  /* compress: fewer than 5 bytes — reference never starts (deflate.py:429-432) */
    HDLZ_ST_BAD_BTYPE = 2,    /* "Bad method" (deflate.py:718-721)                                          */
    HDLZ_ST_BAD_CODE = 3,     /* "Invalid data" / "invalid token" / "< 1 bits" (deflate.py:1140,1439,1560)  */
    HDLZ_ST_DIST_TOO_FAR = 4, /* "distance too big" (deflate.py:1506-1508)                                  */
    HDLZ_ST_TRUNCATED = 5,    /* "NO EOF!" (deflate.py:1535-1539)                                           */
    HDLZ_ST_OUT_OVERFLOW = 6, /* output does not fit out_cap / out_stride (reference: ring back-pressure, deflate.py:1531,1597) */
    HDLZ_ST_BAD_STORED = 7,   /* stored block LEN != ~NLEN (reference does not check; zlib does)            */
    HDLZ_ST_BAD_HEADER = 8,   /* zlib CMF/FLG invalid — only with HDLZ_F_VERIFY_HEADER (reference skips it, deflate.py:644,665-676) */
    HDLZ_ST_BAD_ADLER = 9,    /* Adler-32 mismatch — only with HDLZ_F_VERIFY_ADLER (reference never checks, deflate.py:1535)       */
    HDLZ_ST_BAD_CRC = 10,     /* gzip CRC-32 / ISIZE mismatch — only with HDLZ_F_GZIP | HDLZ_F_VERIFY_ADLER                        */
    HDLZ_ST_NO_CODE = 11      /* compress with hdlz_set_tree: the stream holds a symbol the tree has no code for                   */
} hdlz_status;

/* decompress flags */
#define HDLZ_F_VERIFY_HEADER 1u /* zlib: check CMF/FLG (the reference skips the two bytes, deflate.py:644)           */
#define HDLZ_F_VERIFY_ADLER 2u  /* check the container checksum: Adler-32 (zlib) or CRC-32 + ISIZE (gzip)             */
#define HDLZ_F_RAW 4u           /* input is a bare RFC 1951 stream: no header, no trailer                              */
#define HDLZ_F_GZIP 8u          /* input is one RFC 1952 (gzip) member; header always checked, optional fields skipped */
#define HDLZ_F_PERSIST_TABLES 16u /* batches of dynamic-Huffman streams: keep the per-stream decode tables in the L2
                                  * (sets the device's persisting-L2 carve-out until the next call without the flag) */

/* Container written / read around the deflate body.  The reference knows zlib only (header bytes
 * deflate.py:753-757, Adler-32 :788-814); raw and gzip wrap the SAME body (README.md:2 "(g)zip / zlib"). */
typedef enum hdlz_container {
    HDLZ_CONTAINER_ZLIB = 0,   /* 78 9C | body | Adler-32 (big-endian)                       — the default      */
    HDLZ_CONTAINER_RAW = 1,    /* body only                                                                     */
    HDLZ_CONTAINER_GZIP = 2    /* 1F 8B 08 00 00000000 00 FF | body | CRC-32 | ISIZE (little-endian)            */
} hdlz_container;

typedef struct hdlz_ctx hdlz_ctx;

/* ---- library ------------------------------------------------------------ */
int hdlz_version(void);
const char *hdlz_last_error(void);
const char *hdlz_status_name(uint32_t status);
int hdlz_device_count(void);

/* One context per (process, GPU).  Replaces instantiating the block,
 * `dut = deflate(i_mode, ..., clk, reset)` (deflate.py:219-221; test_deflate.py:314):
 * owns the scratch the reference keeps in iram/oram/leaves (deflate.py:229-230,277-280). */
int hdlz_create(int device, hdlz_ctx **ctx);
int hdlz_destroy(hdlz_ctx *ctx);

/* Worst-case compressed size of `len` input bytes: 2 + ceil((3 + 9*len + 7)/8) + 4,
 * rounded up to 16 (all-9-bit literals; CSTATIC framing deflate.py:746-814). */
uint32_t hdlz_compress_bound(uint32_t len);

/* The reference's module switch MATCH10 (deflate.py:34-35): non-zero (the default, and what the
 * BASELINE configs use) lets SEARCHF grow a match to 10 bytes, zero stops it at 5 (deflate.py:913-924).
 * Applies to every later compress call of the context; output is bit-identical to deflate.py built
 * with the same setting (FAST = True, CWINDOW = 32). */
int hdlz_set_match10(hdlz_ctx *ctx, int match10);
int hdlz_get_match10(hdlz_ctx *ctx);

/* The reference's module switch FAST (deflate.py:36-37, 56-59): non-zero (the default, and what the BASELINE
 * configs use) is the 32-byte window with the parallel `matcher3` search; zero selects the non-FAST engine —
 * CWINDOW = 256, SEARCH / SEARCH10 walking back from the nearest position (deflate.py:996-1062), distance
 * codes up to 15 (`outcarry`, :875-880).  Output is bit-identical to deflate.py built with the same FAST and
 * MATCH10 settings.  A slower second mode: the search is direct, not the mask formulation of the FAST kernel. */
int hdlz_set_fast(hdlz_ctx *ctx, int fast);
int hdlz_get_fast(hdlz_ctx *ctx);

/* Container of every later compress call of the context (hdlz_container; default zlib, the
 * reference's).  The deflate body is the same bits in all three.  hdlz_compress_bound_ex is the
 * slot size a stream of `len` bytes needs in the given container (hdlz_compress_bound = zlib). */
int hdlz_set_container(hdlz_ctx *ctx, int container);
int hdlz_get_container(hdlz_ctx *ctx);
uint32_t hdlz_compress_bound_ex(uint32_t len, int container);

/* The code the compressor writes with.  The reference always codes with the fixed tree of RFC 1951
 * (out_codes, deflate.py:112-149; STATIC :1064-1076) and names "a dedicated pre-computed Huffman tree"
 * for data with few byte values as the next step (README.md:43-45).  hdlz_set_tree installs one:
 * code lengths (0 = unused, <= 15) of the 286 literal/length and 30 distance symbols; every later
 * compress call of the context writes ONE dynamic-Huffman block (BTYPE = 10) per stream that opens
 * with the description of this code and holds the same tokens deflate.py's parse produces (SEARCH /
 * SEARCHF / DISTANCE, deflate.py:836-1016), coded with it.  Symbol 256 needs a code; a stream that
 * needs a symbol without one ends with HDLZ_ST_NO_CODE.  Lengths must not be over-subscribed and may be
 * incomplete only as a single one-bit code (what inflaters accept).  NULL, NULL returns to the fixed code.
 * hdlz_train_tree builds the lengths from a batch on the device: the compress kernel counts the symbols
 * of its parse (no output), the counts (+1 for every symbol the parse can produce, so that a later batch
 * cannot meet a missing code) become optimal length-limited code lengths, which are installed.
 * hdlz_get_tree returns 1 and copies the lengths when a tree is installed, 0 otherwise.
 * hdlz_compress_bound_tree: slot size a stream of `len` bytes needs with the installed code.
 * FAST compressor (CWINDOW = 32) only; whole-stream and batch calls (not hdlz_cstream_*).  Setting or
 * training a tree waits for the device to go idle first (earlier launches may still read the old tables). */
int hdlz_set_tree(hdlz_ctx *ctx, const uint8_t *lit_len /* [286] */, const uint8_t *dist_len /* [30] */);
int hdlz_get_tree(hdlz_ctx *ctx, uint8_t *lit_len, uint8_t *dist_len);
int hdlz_train_tree(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                    uint32_t uniform_len, uint64_t n, void *stream);
uint32_t hdlz_compress_bound_tree(hdlz_ctx *ctx, uint32_t len);
/* One stream coded with ITS OWN tree — the dynamic-tree compression the reference leaves to future work
 * (README.md:65 "HDL-Deflate compressed output is always using a static tree"): the symbols of the parse of the
 * stream's 2 KiB blocks are counted on the GPU, the optimal length-limited code is built and the stream is coded
 * with it in one BTYPE = 10 block (tokens: deflate.py's parse of the whole stream, as always).  Streams shorter than
 * 2 KiB use the fixed code.  The context's own tree setting is restored before the call returns.  Arguments and
 * status as hdlz_compress_stream.  out_cap >= hdlz_compress_bound(len) + 8192 always suffices (code description, the
 * +1 counts, and up to 2047 tail bytes outside the counted blocks at 15 bits each); a smaller out_cap that the stream
 * does not fit gives HDLZ_ST_OUT_OVERFLOW, never an overrun. */
int hdlz_compress_stream_dyn(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                             uint32_t *out_len, uint32_t *status);
/* Device-independent helpers for an application that gathers its own statistics: optimal code lengths
 * (<= max_bits) for `n` symbol counts (0 = unused symbol, gets length 0), and the bytes every stream of a
 * context with these lengths starts with (container header, BFINAL / BTYPE = 10, RFC 1951 3.2.7 code
 * description; *out_bits of them, the last byte zero-padded). */
int hdlz_tree_lengths(const uint64_t *count, int n, int max_bits, uint8_t *len);
int hdlz_tree_header(const uint8_t *lit_len, const uint8_t *dist_len, int container, uint8_t *out, uint32_t out_cap,
                     uint32_t *out_bits);

/* ---- compress: STARTC job (deflate.py:618-633; CSTATIC/SEARCH/SEARCHF/DISTANCE/CHECKSUM :734-1016) ----
 * Block i = d_in[i*in_stride .. +len_i), len_i = d_in_len ? d_in_len[i] : uniform_len
 * (the reference's `isize + 1`, deflate.py:605).  Writes one zlib stream (78 9C, one
 * fixed-Huffman block, Adler-32) to d_out[i*out_stride ..], its length to d_out_len[i]
 * (the reference's o_oprogress at o_done) and the outcome to d_status[i] (may be NULL).
 * Output is bit-identical to deflate.py with FAST=MATCH10=True, CWINDOW=32.
 * out_stride must be >= hdlz_compress_bound(max len); bytes of the slot past
 * out_len up to the next multiple of 4 are zeroed, the rest is left untouched. */
int hdlz_compress_batch(hdlz_ctx *ctx, const uint8_t *d_in, uint64_t in_stride, const uint32_t *d_in_len,
                        uint32_t uniform_len, uint8_t *d_out, uint64_t out_stride, uint32_t *d_out_len,
                        uint32_t *d_status, uint64_t n_blocks, void *stream);

/* ---- decompress: STARTD job (deflate.py:635-651; HEADER..COPY :656-732, :1084-1659) ----
 * Stream i = d_in[off_i .. +d_in_len[i]) with off_i = d_in_off ? d_in_off[i] : i*in_stride;
 * zlib-wrapped deflate (stored, fixed and dynamic blocks, 32 KiB window).  Plain bytes go to
 * d_out[i*out_stride .. ] (at most out_cap), count to d_out_len[i], outcome to d_status[i].
 * Output is byte-identical to zlib's inflate. */
int hdlz_decompress_batch(hdlz_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t in_stride,
                          const uint32_t *d_in_len, uint8_t *d_out, uint64_t out_stride, uint32_t out_cap,
                          uint32_t *d_out_len, uint32_t *d_status, uint64_t n_streams, uint32_t flags,
                          void *stream);

/* ---- host-buffer forms (what a reference-side FFI stub calls) -------------
 * Same arguments with HOST pointers; H2D copy, kernel, D2H copy, synchronise.
 * These replace the WRITE... / START / READ... sequence of the port protocol
 * (deflate.py:599-605; test_deflate.py:120-183, 200-274) for whole buffers. */
int hdlz_compress_host(hdlz_ctx *ctx, const uint8_t *in, uint64_t in_stride, const uint32_t *in_len,
                       uint32_t uniform_len, uint8_t *out, uint64_t out_stride, uint32_t *out_len,
                       uint32_t *status, uint64_t n_blocks);
int hdlz_decompress_host(hdlz_ctx *ctx, const uint8_t *in, const uint64_t *in_off, uint64_t in_stride,
                         const uint32_t *in_len, uint8_t *out, uint64_t out_stride, uint32_t out_cap,
                         uint32_t *out_len, uint32_t *status, uint64_t n_streams, uint32_t flags);

/* ---- packed stream layout -------------------------------------------------
 * hdlz_pack_batch: device, asynchronous.  Packs the fixed-stride slots of a compress batch
 * contiguously: stream i goes to d_packed[d_off[i] .. +d_len[i]) with d_off[i] the exclusive prefix
 * sum of the lengths rounded up to 4; *d_total (device) receives the packed size.  d_off feeds
 * hdlz_decompress_batch directly.  d_packed needs sum(round4(len)) bytes (<= n * stride).
 * hdlz_compress_host_packed: hdlz_compress_host, but `out` (capacity out_cap bytes) receives the
 * streams packed and out_off[i] their offsets; *out_total the bytes used.  Only real stream bytes
 * are copied back.  The reference has no counterpart (its oram holds one stream, deflate.py:230). */
int hdlz_pack_batch(hdlz_ctx *ctx, const uint8_t *d_slots, uint64_t stride, const uint32_t *d_len,
                    uint8_t *d_packed, uint64_t *d_off, uint64_t *d_total, uint64_t n_streams, void *stream);
int hdlz_compress_host_packed(hdlz_ctx *ctx, const uint8_t *in, uint64_t in_stride, const uint32_t *in_len,
                              uint32_t uniform_len, uint8_t *out, uint64_t out_cap, uint64_t *out_off,
                              uint32_t *out_len, uint32_t *status, uint64_t n_blocks, uint64_t *out_total);

/* One stream of any length < 2^24 (LMAX): exactly what one STARTC / STARTD job of the
 * port protocol does.  `status` receives the hdlz_status.  A compress stream of HDLZ_LONG_STREAM bytes or
 * more is spread over the whole GPU, a tile of 1024 positions per warp (fixed or installed tree, FAST, zlib /
 * raw container; otherwise one warp works through it): the parse position, the bit cursor and the Adler sums —
 * what the reference's FSM carries from byte to byte — cross the tile borders by look-back between the
 * warps, and the bytes are the same.  hdlz_compress_batch does the same for a batch of fewer than 32 x SMs
 * streams of one length (d_in_len == NULL) of HDLZ_LONG_STREAM bytes or more. */
int hdlz_compress_stream(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                         uint32_t *out_len, uint32_t *status);
int hdlz_decompress_stream(hdlz_ctx *ctx, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                           uint32_t *out_len, uint32_t *status, uint32_t flags);

/* ---- one stream fed in pieces --------------------------------------------
 * The port protocol streams: the host WRITEs input while the engine runs, and the engine stalls at
 * `di >= isize - 10` until more bytes arrive or IDLE says there are none (deflate.py:459-461, 768; flow
 * control of the test bench, test_deflate.py:159, 250).  hdlz_cstream_* is that for the compressor: feed any
 * number of pieces of any size, collect the stream bytes as they are completed; the engine carries the bit
 * cursor and partial output word (`do` / `doo` / `ob1`), the parse position, the last 32 bytes and both
 * Adler sums from call to call (state lives on the device).  The concatenated output is bit-identical to
 * one hdlz_compress_stream call over the whole input (and so to deflate.py).  zlib and raw containers.
 *   feed:   consumes all of `in`; writes the bytes completed so far to out (at most out_cap; size it
 *           hdlz_compress_bound(len + 2048)), *out_len their count, *in_progress the input position up to
 *           which the stream is encoded (the reference's o_iprogress).
 *   finish: no more input (the reference's IDLE after START): the rest of the stream, EOB, Adler-32.
 *           *status as hdlz_compress_stream (HDLZ_ST_SHORT_INPUT for fewer than 5 bytes in total). */
/* (a piece that completes 64 tiles or more is encoded by the whole GPU, like hdlz_compress_stream of a long input;
 * smaller pieces by one warp — the bytes are the same either way) */
typedef struct hdlz_cstream hdlz_cstream;
int hdlz_cstream_begin(hdlz_ctx *ctx, hdlz_cstream **stream);
int hdlz_cstream_feed(hdlz_cstream *stream, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                      uint32_t *out_len, uint32_t *in_progress);
int hdlz_cstream_finish(hdlz_cstream *stream, uint8_t *out, uint32_t out_cap, uint32_t *out_len, uint32_t *status);
int hdlz_cstream_end(hdlz_cstream *stream);

/* The decompressor's counterpart.  The reference inflates while the host is still writing the stream: the FSM
 * waits at `di >= isize - 4` until more input arrives or the host goes IDLE (deflate.py:1529), its output ring
 * fills as it goes and the host reads bytes as o_oprogress advances (deflate.py:1531, 1597; test bench
 * test_deflate.py:136-194).  hdlz_dstream_* is that: feed the stream in pieces of any size; every call decodes
 * as far as the input received so far allows — the kernel stops before a block header it cannot see whole
 * (320 bytes) or a symbol with fewer than 64 bits left, records bit cursor, output cursor and the position of
 * the current block's header on the device, and the next call carries on from there — and hands back up to
 * out_cap of the output bytes not yet delivered (what does not fit stays on the device: the back-pressure of
 * the reference's output ring).  The bytes of all calls, joined, are what hdlz_decompress_stream gives for the
 * whole stream.  `flags` as hdlz_decompress_batch (containers, verification; checksums are checked by the
 * closing call).
 *   begin   max_out = capacity for the whole output (< 2^HDLZ_LMAX)
 *   feed    out <- up to out_cap new output bytes, *out_len their count, *in_progress the input byte position
 *           the decoder has reached (o_iprogress)
 *   finish  no more input: decodes to the end.  Delivers up to out_cap bytes; *remaining = bytes still to be
 *           fetched (call finish again with more room), *status as hdlz_decompress_stream. */
typedef struct hdlz_dstream hdlz_dstream;
int hdlz_dstream_begin(hdlz_ctx *ctx, uint32_t max_out, uint32_t flags, hdlz_dstream **stream);
int hdlz_dstream_feed(hdlz_dstream *stream, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t out_cap,
                      uint32_t *out_len, uint32_t *in_progress);
int hdlz_dstream_finish(hdlz_dstream *stream, uint8_t *out, uint32_t out_cap, uint32_t *out_len, uint32_t *remaining,
                        uint32_t *status);
int hdlz_dstream_end(hdlz_dstream *stream);

/* ---- device memory helpers (for hosts without their own CUDA allocator) --- */
int hdlz_dev_alloc(hdlz_ctx *ctx, size_t bytes, void **d_ptr);
int hdlz_dev_free(hdlz_ctx *ctx, void *d_ptr);
int hdlz_host_alloc_pinned(hdlz_ctx *ctx, size_t bytes, void **h_ptr);
int hdlz_host_free_pinned(hdlz_ctx *ctx, void *h_ptr);
int hdlz_copy_h2d(hdlz_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int hdlz_copy_d2h(hdlz_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);
int hdlz_stream_sync(hdlz_ctx *ctx, void *stream);

/* ---- measurement support --------------------------------------------------
 * Synthetic "random+repeat" blocks of BASELINE config 2: block b is a
 * pure function of (seed, first_block + b); see hdl-deflate_b200/workload.py for the
 * bit-identical CPU definition.  Kernel launches issued by this library since
 * hdlz_create (for the bench's gpu_launches count). */
int hdlz_generate_blocks(hdlz_ctx *ctx, uint8_t *d_out, uint64_t stride, uint32_t len, uint64_t n_blocks,
                         uint64_t seed, uint64_t first_block, void *stream);
uint64_t hdlz_launch_count(hdlz_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* HDLZ_H */
