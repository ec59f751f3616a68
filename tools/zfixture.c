/* zfixture.c — measurement plumbing for bench.py: HOST zlib streams of many equal-sized rows on all
 * cores (fixture generator of BASELINE configs[2] "zlib Z_FIXED streams" and configs[3] "zlib level-6
 * streams").  Python's zlib costs ~100 us of interpreter time per 2 KiB row, this does 2^20 rows in a
 * few seconds.  Not part of the product (nothing under hdl-deflate_b200/ links it) and not the oracle:
 * it only produces INPUT for the CUDA decompressor.  Built on demand: gcc -O2 -shared -fPIC -lz -lpthread. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
    const uint8_t *in; uint64_t lo, hi; uint32_t L; int level, strategy;
    uint8_t *tmp; uint64_t used; uint32_t *len; uint64_t *off; int err;
    uint8_t *out; uint64_t base;
} zjob_t;

static void *zwork(void *a)
{
    zjob_t *j = (zjob_t *)a;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, j->level, Z_DEFLATED, 15, 8, j->strategy) != Z_OK) { j->err = 1; return NULL; }
    const uint64_t bound = (deflateBound(&zs, j->L) + 3u) & ~3ull;
    j->tmp = (uint8_t *)malloc((j->hi - j->lo) * bound + 16);
    if (!j->tmp) { j->err = 1; deflateEnd(&zs); return NULL; }
    uint64_t pos = 0;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        deflateReset(&zs);
        zs.next_in = (Bytef *)(j->in + i * j->L); zs.avail_in = j->L;
        zs.next_out = j->tmp + pos; zs.avail_out = (uInt)bound;
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { j->err = 1; break; }
        const uint32_t n = (uint32_t)zs.total_out;
        j->len[i] = n;
        j->off[i] = pos;                       /* thread-local; rebased after the join */
        const uint64_t next = pos + ((n + 3u) & ~3ull);
        memset(j->tmp + pos + n, 0, next - (pos + n));
        pos = next;
    }
    j->used = pos;
    deflateEnd(&zs);
    return NULL;
}

static void *zcopy(void *a)
{
    zjob_t *j = (zjob_t *)a;
    memcpy(j->out + j->base, j->tmp, j->used);
    for (uint64_t i = j->lo; i < j->hi; i++) j->off[i] += j->base;
    free(j->tmp);
    return NULL;
}

/* rows in[i*L .. +L) -> zlib streams packed into out (starts 4-byte aligned, row order), off[i], len[i];
 * *total = bytes used.  Returns 0, 1 on a zlib / allocation failure, 2 if out_cap is too small. */
int zfix_deflate_packed(const uint8_t *in, uint64_t n, uint32_t L, int level, int strategy, int threads,
                        uint8_t *out, uint64_t out_cap, uint64_t *off, uint32_t *len, uint64_t *total)
{
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    zjob_t *jobs = (zjob_t *)calloc(threads, sizeof(zjob_t));
    for (int t = 0; t < threads; t++) {
        zjob_t j = {in, n * t / threads, n * (t + 1) / threads, L, level, strategy, NULL, 0, len, off, 0, out, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, zwork, &jobs[t]);
    }
    int err = 0;
    uint64_t base = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        err |= jobs[t].err;
        jobs[t].base = base;
        base += jobs[t].used;
    }
    if (!err && base > out_cap) err = 2;
    if (err) {
        for (int t = 0; t < threads; t++) free(jobs[t].tmp);
    } else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, zcopy, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    *total = base;
    free(th); free(jobs);
    return err;
}
