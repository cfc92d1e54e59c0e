/* itx_bgzf.c -- BGZF container handling on the host: block table + multi-threaded raw-deflate
 * inflate straight into pinned staging memory.
 *
 * Restates the reading side of cussamtools/bgzf.c: check_header 401-411 (gzip magic, FEXTRA, XLEN 6,
 * 'B''C' subfield, BSIZE at byte 16), bgzf_read_block 471-521, inflate_block 367-397 (windowBits -15,
 * no CRC comparison), bgzf_read 524-565 (an empty block, a bad header or a short read ends the
 * stream: bam_read1 then fails and the reference's loop stops silently, generic.c:745).
 * Unlike the reference, blocks are inflated out of order by a pool of threads; every block's
 * uncompressed offset is known beforehand from the ISIZE footers.
 */
#define _GNU_SOURCE
#include "itx_internal.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

int itx_bgzf_header_ok(const uint8_t *h) {
    return h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
}

/* First BGZF block boundary in buf[0, n): an offset whose header checks out and whose BSIZE leads to another good header --
 * twice over when the bytes are there -- or exactly to the end of the file (at_eof).  What a reader that is dropped into the
 * middle of a file does instead of consulting a .bai (the sharded scan: every rank but the first).  Returns -1 if there is none. */
int64_t itx_bgzf_find_block(const uint8_t *buf, uint64_t n, int at_eof) {
    for (uint64_t i = 0; i + 18 <= n; i++) {
        if (buf[i] != 31 || !itx_bgzf_header_ok(buf + i)) continue;
        uint64_t o = i; int good = 0;
        for (int hop = 0; hop < 3; hop++) {
            if (o + 18 > n) { good = hop > 0 || (at_eof && o == n); break; }
            if (!itx_bgzf_header_ok(buf + o)) { good = 0; break; }
            const uint32_t bsize = ((uint32_t)buf[o + 16] | (uint32_t)buf[o + 17] << 8) + 1;
            if (bsize < 26) { good = 0; break; }
            o += bsize; good = 1;
            if (at_eof && o == n) break;
        }
        if (good) return (int64_t)i;
    }
    return -1;
}

int itx_bgzf_scan(const uint8_t *file, uint64_t len, itx_bgzf_block **blocks, uint64_t *n_blocks, uint64_t *total_u, char err[ITX_ERRLEN]) {
    uint64_t cap = len / 16384 + 64, n = 0, off = 0, u = 0;
    itx_bgzf_block *b = (itx_bgzf_block *)malloc(sizeof(itx_bgzf_block) * cap);
    if (!b) { snprintf(err, ITX_ERRLEN, "out of memory"); return ITX_ENOMEM; }
    while (off + 18 <= len) {
        const uint8_t *h = file + off;
        if (!itx_bgzf_header_ok(h)) break;
        uint32_t bsize = ((uint32_t)h[16] | (uint32_t)h[17] << 8) + 1;
        if (bsize < 26 || off + bsize > len) break;                    /* short read: the stream ends here */
        uint32_t isize; memcpy(&isize, h + bsize - 4, 4);
        if (isize == 0) break;                                         /* empty block = end of data */
        if (isize > 65536) break;
        if (n == cap) { cap *= 2; b = (itx_bgzf_block *)realloc(b, sizeof(itx_bgzf_block) * cap); }
        b[n].coff = off; b[n].csize = bsize; b[n].isize = isize; b[n].uoff = u;
        n++; u += isize; off += bsize;
    }
    *blocks = b; *n_blocks = n; *total_u = u;
    return ITX_OK;
}

/* ------------------------------------------------------------------ thread pool */
typedef struct {
    pthread_mutex_t mu; pthread_cond_t cv_work, cv_done;
    pthread_t *th; int nth; int generation, pending, stop;
    /* the current job */
    const uint8_t *file; const itx_bgzf_block *blk; uint64_t b0, b1; uint8_t *dst; uint64_t next; int failed;
    int mode;                            /* 0: inflate blocks [b0,b1)   1: copy bytes [b0,b1) of file to dst in 1 MiB pieces   2: the same with pread(fd) */
    int fd;
    double busy_max;
} pool_t;
static pool_t g_pool; static int g_pool_init = 0; static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;

static double mono_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

static void *worker(void *arg) {
    pool_t *P = (pool_t *)arg;
    z_stream zs; memset(&zs, 0, sizeof zs);
    int zinit = inflateInit2(&zs, -15) == Z_OK;
    int seen = 0;
    for (;;) {
        pthread_mutex_lock(&P->mu);
        while (!P->stop && P->generation == seen) pthread_cond_wait(&P->cv_work, &P->mu);
        if (P->stop) { pthread_mutex_unlock(&P->mu); break; }
        seen = P->generation;
        pthread_mutex_unlock(&P->mu);
        double t0 = mono_ms(); int failed = !zinit;
        if (P->mode == 1 || P->mode == 2) {
            failed = 0;
            for (;;) {
                uint64_t o = __atomic_fetch_add(&P->next, (uint64_t)1 << 20, __ATOMIC_RELAXED);
                if (o >= P->b1) break;
                uint64_t e = o + ((uint64_t)1 << 20) < P->b1 ? o + ((uint64_t)1 << 20) : P->b1;
                if (P->mode == 1) memcpy(P->dst + (o - P->b0), P->file + o, (size_t)(e - o));
                else {
                    while (o < e) {
                        ssize_t r = pread(P->fd, P->dst + (o - P->b0), (size_t)(e - o), (off_t)o);
                        if (r <= 0) { failed = 1; break; }
                        o += (uint64_t)r;
                    }
                }
            }
        } else
        for (;;) {
            uint64_t i = __atomic_fetch_add(&P->next, 4, __ATOMIC_RELAXED);   /* four blocks per grab */
            if (i >= P->b1) break;
            uint64_t e = i + 4 < P->b1 ? i + 4 : P->b1;
            for (; i < e && !failed; i++) {
                const itx_bgzf_block *B = &P->blk[i];
                inflateReset2(&zs, -15);
                zs.next_in = (Bytef *)(P->file + B->coff + 18); zs.avail_in = B->csize - 16;
                zs.next_out = P->dst + (B->uoff - P->blk[P->b0].uoff); zs.avail_out = B->isize;
                int st = inflate(&zs, Z_FINISH);
                if (st != Z_STREAM_END || zs.total_out != B->isize) failed = 1;
            }
        }
        double dt = mono_ms() - t0;
        pthread_mutex_lock(&P->mu);
        if (failed) P->failed = 1;
        if (dt > P->busy_max) P->busy_max = dt;
        if (--P->pending == 0) pthread_cond_signal(&P->cv_done);
        pthread_mutex_unlock(&P->mu);
    }
    if (zinit) inflateEnd(&zs);
    return NULL;
}

static void pool_start(int nth) {
    pool_t *P = &g_pool;
    memset(P, 0, sizeof *P);
    pthread_mutex_init(&P->mu, NULL); pthread_cond_init(&P->cv_work, NULL); pthread_cond_init(&P->cv_done, NULL);
    P->th = (pthread_t *)calloc((size_t)nth, sizeof(pthread_t));
    for (int i = 0; i < nth; i++) if (pthread_create(&P->th[P->nth], NULL, worker, P) == 0) P->nth++;
}
static void pool_stop(void) {
    pool_t *P = &g_pool;
    pthread_mutex_lock(&P->mu); P->stop = 1; pthread_cond_broadcast(&P->cv_work); pthread_mutex_unlock(&P->mu);
    for (int i = 0; i < P->nth; i++) pthread_join(P->th[i], NULL);
    free(P->th); P->th = NULL; P->nth = 0;
}

static int pool_run(int mode, int fd, const uint8_t *file, const itx_bgzf_block *blocks, uint64_t b0, uint64_t b1, uint8_t *dst, int nth, double *busy_ms);
int itx_bgzf_inflate_range(const uint8_t *file, const itx_bgzf_block *blocks, uint64_t b0, uint64_t b1, uint8_t *dst, int nth, double *busy_ms) {
    return pool_run(0, -1, file, blocks, b0, b1, dst, nth, busy_ms);
}
/* dst[0 .. o1-o0) = file[o0 .. o1) with nth threads (staging a compressed window into pinned memory) */
int itx_parallel_copy(const uint8_t *file, uint64_t o0, uint64_t o1, uint8_t *dst, int nth) {
    return pool_run(1, -1, file, NULL, o0, o1, dst, nth, NULL);
}
/* the same straight from a file descriptor: no mapping, so no page faults -- the kernel copies out of the page cache */
int itx_parallel_pread(int fd, uint64_t o0, uint64_t o1, uint8_t *dst, int nth) {
    return pool_run(2, fd, NULL, NULL, o0, o1, dst, nth, NULL);
}
static int pool_run(int mode, int fd, const uint8_t *file, const itx_bgzf_block *blocks, uint64_t b0, uint64_t b1, uint8_t *dst, int nth, double *busy_ms) {
    if (b1 <= b0) return ITX_OK;
    pthread_mutex_lock(&g_pool_mu);
    if (g_pool_init && g_pool.nth != nth) { pool_stop(); g_pool_init = 0; }
    if (!g_pool_init) { pool_start(nth); g_pool_init = 1; }
    pool_t *P = &g_pool;
    if (P->nth == 0) { pthread_mutex_unlock(&g_pool_mu); return ITX_ENOMEM; }
    pthread_mutex_lock(&P->mu);
    P->file = file; P->blk = blocks; P->b0 = b0; P->b1 = b1; P->dst = dst; P->next = b0; P->failed = 0; P->busy_max = 0; P->mode = mode;
    P->fd = fd; P->pending = P->nth; P->generation++;
    pthread_cond_broadcast(&P->cv_work);
    while (P->pending) pthread_cond_wait(&P->cv_done, &P->mu);
    int failed = P->failed; if (busy_ms) *busy_ms = P->busy_max;
    pthread_mutex_unlock(&P->mu);
    pthread_mutex_unlock(&g_pool_mu);
    return failed ? ITX_EFORMAT : ITX_OK;
}
