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
    hdlz_destroy(ctx);
    free(comp);
    free(back);
    return rc;
}
