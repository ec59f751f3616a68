/* A C client of include/hdlz.h: what a cgo / JNI / FFI stub binds.  Built and run by
 * tests/test_abi.py (no GPU needed for the calls it makes; on a GPU box it also runs one job). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hdlz.h"

int main(void)
{
    printf("version %06x\n", hdlz_version());
    printf("bound2048 %u %u %u\n", hdlz_compress_bound(2048), hdlz_compress_bound_ex(2048, HDLZ_CONTAINER_RAW),
           hdlz_compress_bound_ex(2048, HDLZ_CONTAINER_GZIP));
    printf("status4 %s\n", hdlz_status_name(HDLZ_ST_DIST_TOO_FAR));
    /* device-independent helpers of the tree-coded mode: code lengths from counts, the stream prefix */
    {
        uint64_t count[286] = {0};
        uint8_t lit[286], dist[30] = {0}, hdr[400];
        uint32_t bits = 0;
        const char *alphabet = "ACGT\n";
        for (const char *c = alphabet; *c; ++c) count[(unsigned char)*c] = 1000;
        for (int s = 256; s < 265; ++s) count[s] = 10;
        int trc = hdlz_tree_lengths(count, 286, 15, lit);
        for (int c = 0; c < 10; ++c) dist[c] = c < 6 ? 3 : 4;          /* 6/8 + 4/16 = 1: a complete code */
        trc |= hdlz_tree_header(lit, dist, HDLZ_CONTAINER_ZLIB, hdr, sizeof hdr, &bits);
        printf("tree %d lenA %u lenEOB %u prefix_bits %u head %02x%02x\n", trc, lit['A'], lit[256], bits, hdr[0], hdr[1]);
    }
    int n = hdlz_device_count();
    printf("devices %d\n", n);
    hdlz_ctx *ctx = NULL;
    int rc = hdlz_create(0, &ctx);
    if (rc != HDLZ_SUCCESS) {
        printf("create %d %s\n", rc, hdlz_last_error());     /* no CPU path: the library says so */
        return n == 0 ? 0 : 1;
    }
    const char *text = "   Hello World! 0        Hello World! 1        Hello World! 2     ";
    uint32_t len = (uint32_t)strlen(text), cap = hdlz_compress_bound(len), clen = 0, st = 0, blen = 0;
    uint8_t *comp = malloc(cap), *back = malloc(len);
    rc = hdlz_compress_stream(ctx, (const uint8_t *)text, len, comp, cap, &clen, &st);
    printf("compress %d status %u bytes %u head %02x%02x\n", rc, st, clen, comp[0], comp[1]);
    rc |= hdlz_decompress_stream(ctx, comp, clen, back, len, &blen, &st, HDLZ_F_VERIFY_HEADER | HDLZ_F_VERIFY_ADLER);
    printf("decompress %d status %u bytes %u same %d\n", rc, st, blen, blen == len && memcmp(back, text, len) == 0);
    /* the same stream inflated in pieces (hdlz_dstream_*): 7 bytes of input per call */
    {
        hdlz_dstream *ds = NULL;
        uint8_t piece[256];
        uint32_t got = 0, produced = 0, prog = 0, remaining = 0;
        memset(back, 0, len);
        rc |= hdlz_dstream_begin(ctx, len, HDLZ_F_VERIFY_ADLER, &ds);
        for (uint32_t pos = 0; pos < clen && rc == HDLZ_SUCCESS; pos += 7) {
            const uint32_t k = clen - pos < 7 ? clen - pos : 7;
            rc |= hdlz_dstream_feed(ds, comp + pos, k, piece, sizeof piece, &produced, &prog);
            memcpy(back + got, piece, produced);
            got += produced;
        }
        rc |= hdlz_dstream_finish(ds, piece, sizeof piece, &produced, &remaining, &st);
        memcpy(back + got, piece, produced);
        got += produced;
        hdlz_dstream_end(ds);
        printf("dstream %d status %u bytes %u remaining %u same %d\n", rc, st, got, remaining,
               got == len && memcmp(back, text, len) == 0);
    }
    hdlz_destroy(ctx);
    free(comp);
    free(back);
    return rc;
}
