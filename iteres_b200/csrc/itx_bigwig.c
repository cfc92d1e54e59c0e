/* itx_bigwig.c -- wiggle -> bigWig, the output step that follows the hot path in `iteres stat` and `iteres cpgstat`
 * (stat.c:157-158, cpgstat.c:75: bigWigFileCreate(wig, rep_size_file, 256, 1024, 0, 1, bigWig)).
 *
 * A from-scratch writer of the bigWig container that reproduces, byte for byte, what the reference's vendored Kent
 * code writes for the wiggle files iteres makes (fixedStep sections, one per subfamily with a known consensus
 * length).  What is restated, file:line in the reference tree:
 *   cuskent/bwgCreate.c   186-264 fixedStep sections of <= itemsPerSlot values; 45-136 section record (24-byte header +
 *                         float values, zlib compress()); 138-151 section order; 583-625 chromosome ids; 627-689 average
 *                         resolution; 788-1015 zoom-level choice, file layout, total summary, header patch-up
 *   cuskent/bbiWrite.c    368-422 folding a range into the running summary (integer / float truncation as in C);
 *                         435-446 summary -> coarser summary; 478-536 compressed summary blocks + their index
 *   cuskent/cirTree.c     36-334 the R tree: leaf elements of itemsPerSlot items, parents of blockSize children, at
 *                         least two levels; index levels top-down, then leaf nodes (short nodes are padded with
 *                         24-byte slots, also in leaf nodes whose slots are 32 bytes)
 *   cuskent/bPlusTree.c   405-577 the chromosome name -> (id, size) B+ tree
 * Host C; nothing here runs on the device.
 */
#define _GNU_SOURCE
#include "itx_internal.h"
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <zlib.h>

#define BW_SIG 0x888FFC26u
#define BPT_SIG 0x78CA8C91u
#define CIR_SIG 0x2468ACE0u
#define BW_VERSION 4

typedef struct {
    char *chrom; uint32_t chrom_id, start, end, step, span; uint16_t count; float *val; uint64_t file_off;
} bw_section;
typedef struct bw_sum { uint32_t chrom_id, start, end, valid; float minv, maxv, sum, sumsq; uint64_t file_off; } bw_sum;
typedef struct { bw_sum *v; size_t n, cap; } bw_sumlist;
typedef struct { char *name; uint32_t id, size; } bw_chrom;

#define W(f, x) fwrite(&(x), sizeof(x), 1, f)
static void zeros(FILE *f, size_t n) { while (n--) fputc(0, f); }

/* ---- summaries (bbiWrite.c:368-446) */
static bw_sum *sum_push(bw_sumlist *L) {
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 1024; L->v = (bw_sum *)realloc(L->v, L->cap * sizeof(bw_sum)); }
    memset(&L->v[L->n], 0, sizeof(bw_sum));
    return &L->v[L->n++];
}
static void sum_add(bw_sumlist *L, uint32_t chrom_id, uint32_t chrom_size, uint32_t start, uint32_t end, uint32_t valid,
                    double minv, double maxv, double sumd, double sumsq, int reduction) {
    bw_sum *s = L->n ? &L->v[L->n - 1] : NULL;
    if (end > chrom_size) end = chrom_size;
    while (start < end) {
        if (!s || s->chrom_id != chrom_id || s->end <= start) {
            uint32_t ns;
            if (!s || s->chrom_id != chrom_id || s->end + (uint32_t)reduction <= start) ns = start; else ns = s->end;
            s = sum_push(L);
            s->chrom_id = chrom_id; s->start = ns; s->end = ns + (uint32_t)reduction;
            if (s->end > chrom_size) s->end = chrom_size;
            s->minv = (float)minv; s->maxv = (float)maxv;
        }
        int lo = (int)(start > s->start ? start : s->start), hi = (int)(end < s->end ? end : s->end);
        int overlap = hi - lo;
        if (overlap <= 0) return;                                   /* the reference aborts here (internalErr); cannot happen for sorted input */
        int item = (int)(end - start);
        double factor = (double)overlap / item;
        s->valid = (uint32_t)(s->valid + factor * valid);
        if (s->minv > minv) s->minv = (float)minv;
        if (s->maxv < maxv) s->maxv = (float)maxv;
        s->sum = (float)(s->sum + factor * sumd);
        s->sumsq = (float)(s->sumsq + factor * sumsq);
        start += (uint32_t)overlap;
    }
}
static void reduce_sections(const bw_section *sec, size_t nsec, const bw_chrom *chroms, int reduction, bw_sumlist *out) {
    out->n = 0;
    for (size_t i = 0; i < nsec; i++) {
        const bw_section *S = &sec[i]; int start = (int)S->start;
        for (uint32_t k = 0; k < S->count; k++) {
            int size = (int)S->span; double val = S->val[k], sum = size * val, sq = sum * val;
            bw_sum *t = out->n ? &out->v[out->n - 1] : NULL;
            const uint32_t a = (uint32_t)start, b = (uint32_t)start + S->span;
            if (t && t->chrom_id == S->chrom_id && a >= t->start && b <= t->end && b <= chroms[S->chrom_id].size && a < b) {
                /* the item lies inside the running summary: the same five updates sum_add makes, its overlap factor being exactly 1 */
                t->valid = (uint32_t)(t->valid + 1.0 * (uint32_t)size);
                if (t->minv > val) t->minv = (float)val;
                if (t->maxv < val) t->maxv = (float)val;
                t->sum = (float)(t->sum + 1.0 * sum);
                t->sumsq = (float)(t->sumsq + 1.0 * sq);
            } else sum_add(out, S->chrom_id, chroms[S->chrom_id].size, a, b, (uint32_t)size, val, val, sum, sq, reduction);
            start += (int)S->step;
        }
    }
}
static void reduce_sums(const bw_sumlist *in, const bw_chrom *chroms, int reduction, bw_sumlist *out) {
    out->n = 0;
    for (size_t i = 0; i < in->n; i++) {
        const bw_sum *s = &in->v[i];
        sum_add(out, s->chrom_id, chroms[s->chrom_id].size, s->start, s->end, s->valid, s->minv, s->maxv, s->sum, s->sumsq, reduction);
    }
}

/* ---- R tree over (chrom id, start, end) keyed items (cirTree.c:36-360) */
typedef struct { uint32_t sc, sb, ec, eb; uint64_t off0, off1; size_t child0, nchild; } rnode;
typedef struct { uint32_t chrom, start, end; uint64_t off; } rkey;
static void rnode_widen(rnode *p, const rnode *e) {
    if (e->sc < p->sc) { p->sc = e->sc; p->sb = e->sb; } else if (e->sc == p->sc && e->sb < p->sb) p->sb = e->sb;
    if (e->ec > p->ec) { p->ec = e->ec; p->eb = e->eb; } else if (e->ec == p->ec && e->eb > p->eb) p->eb = e->eb;
}
static void write_rtree(FILE *f, const rkey *items, uint64_t n, uint32_t block, uint32_t per_slot, uint64_t end_off) {
    /* levels[0] = leaf elements (per_slot items each); levels[k + 1] = parents of `block` children; at least two levels */
    rnode *lev[64]; size_t cnt[64]; int nlev = 0;
    size_t n0 = (size_t)((n + per_slot - 1) / per_slot);
    rnode *a = (rnode *)calloc(n0 ? n0 : 1, sizeof(rnode));
    for (size_t e = 0; e < n0; e++) {
        uint64_t i = (uint64_t)e * per_slot, m = n - i < per_slot ? n - i : per_slot;
        rnode *el = &a[e];
        el->sc = el->ec = items[i].chrom; el->sb = items[i].start; el->eb = items[i].end;
        el->off0 = items[i].off; el->off1 = (i + m < n) ? items[i + m].off : end_off;
        for (uint64_t j = 1; j < m; j++) {
            const rkey *k = &items[i + j];
            if (k->chrom < el->sc) { el->sc = k->chrom; el->sb = k->start; } else if (k->chrom == el->sc && k->start < el->sb) el->sb = k->start;
            if (k->chrom > el->ec) { el->ec = k->chrom; el->eb = k->end; } else if (k->chrom == el->ec && k->end > el->eb) el->eb = k->end;
        }
    }
    lev[0] = a; cnt[0] = n0; nlev = 1;
    while (cnt[nlev - 1] > 1 || nlev < 2) {
        size_t nc = cnt[nlev - 1], np = (nc + block - 1) / block;
        rnode *p = (rnode *)calloc(np ? np : 1, sizeof(rnode));
        for (size_t q = 0; q < np; q++) {
            size_t c0 = q * block, m = nc - c0 < block ? nc - c0 : block;
            p[q] = lev[nlev - 1][c0]; p[q].child0 = c0; p[q].nchild = m;
            for (size_t j = 1; j < m; j++) rnode_widen(&p[q], &lev[nlev - 1][c0 + j]);
        }
        lev[nlev] = p; cnt[nlev] = np; nlev++;
    }
    const rnode *root = &lev[nlev - 1][0];
    uint32_t magic = CIR_SIG, reserved = 0;
    W(f, magic); W(f, block); W(f, n); W(f, root->sc); W(f, root->sb); W(f, root->ec); W(f, root->eb); W(f, end_off); W(f, per_slot); W(f, reserved);
    /* the reference numbers levels from the root (0) down to the leaf elements (nlev - 1) */
    const uint64_t inode = 4 + 24ull * block, lnode = 4 + 32ull * block;
    uint64_t level_off[64], off = (uint64_t)ftello(f);
    for (int L = 0; L < nlev; L++) { level_off[L] = off; off += cnt[nlev - 1 - L] * inode; }
    const int final_level = nlev - 3;
    for (int L = 0; L <= final_level; L++) {
        const rnode *nodes = lev[nlev - 1 - L]; const rnode *kids = lev[nlev - 2 - L];
        uint64_t child_off = level_off[L + 1]; const uint64_t child_size = (L == final_level) ? lnode : inode;
        for (size_t q = 0; q < cnt[nlev - 1 - L]; q++) {
            uint8_t is_leaf = 0, res = 0; uint16_t c = (uint16_t)nodes[q].nchild;
            W(f, is_leaf); W(f, res); W(f, c);
            for (size_t j = 0; j < nodes[q].nchild; j++) {
                const rnode *k = &kids[nodes[q].child0 + j];
                W(f, k->sc); W(f, k->sb); W(f, k->ec); W(f, k->eb); W(f, child_off);
                child_off += child_size;
            }
            for (uint32_t j = c; j < block; j++) zeros(f, 24);
        }
    }
    {   /* leaf nodes: the level above the leaf elements */
        const rnode *nodes = lev[1]; const rnode *kids = lev[0];
        for (size_t q = 0; q < cnt[1]; q++) {
            uint8_t is_leaf = 1, res = 0; uint16_t c = (uint16_t)nodes[q].nchild;
            W(f, is_leaf); W(f, res); W(f, c);
            for (size_t j = 0; j < nodes[q].nchild; j++) {
                const rnode *k = &kids[nodes[q].child0 + j]; uint64_t size = k->off1 - k->off0;
                W(f, k->sc); W(f, k->sb); W(f, k->ec); W(f, k->eb); W(f, k->off0); W(f, size);
            }
            for (uint32_t j = c; j < block; j++) zeros(f, 24);       /* 24, not 32: as the reference pads them */
        }
    }
    for (int L = 0; L < nlev; L++) free(lev[L]);
}

/* ---- B+ tree chromosome name -> (id, size) (bPlusTree.c:405-577) */
static void write_chrom_tree(FILE *f, const bw_chrom *chroms, uint32_t n, uint32_t block, uint32_t key_size) {
    uint32_t magic = BPT_SIG, reserved = 0, val_size = 8; uint64_t count = n;
    W(f, magic); W(f, block); W(f, key_size); W(f, val_size); W(f, count); W(f, reserved); W(f, reserved);
    int levels = 1; { long ic = n; while (ic > (long)block) { ic = (ic + block - 1) / block; levels++; } }
    char *key = (char *)calloc(key_size + 1, 1);
    uint64_t index_off = (uint64_t)ftello(f);
    for (int lev = levels - 1; lev > 0; lev--) {
        long slot_per = 1; for (int i = 0; i < lev; i++) slot_per *= block;
        long node_per = slot_per * block, node_count = ((long)n + node_per - 1) / node_per;
        uint64_t bytes_index = 4 + (uint64_t)block * (key_size + 8), bytes_leaf = 4 + (uint64_t)block * (key_size + val_size);
        uint64_t next_child = index_off + (uint64_t)node_count * bytes_index;
        for (long i = 0; i < (long)n; i += node_per) {
            long c = ((long)n - i + slot_per - 1) / slot_per; if (c > (long)block) c = block;
            uint8_t is_leaf = 0, res = 0; uint16_t c16 = (uint16_t)c;
            W(f, is_leaf); W(f, res); W(f, c16);
            long end_ix = i + node_per; if (end_ix > (long)n) end_ix = n;
            for (long j = i; j < end_ix; j += slot_per) {
                memset(key, 0, key_size); strcpy(key, chroms[j].name);
                fwrite(key, 1, key_size, f); W(f, next_child);
                next_child += lev == 1 ? bytes_leaf : bytes_index;
            }
            for (long j = c; j < (long)block; j++) zeros(f, key_size + 8);
        }
        index_off = (uint64_t)ftello(f);
    }
    for (uint32_t i = 0; i < n; ) {
        uint16_t c = (uint16_t)(n - i > block ? block : n - i); uint8_t is_leaf = 1, res = 0;
        W(f, is_leaf); W(f, res); W(f, c);
        for (uint16_t j = 0; j < c; j++) {
            memset(key, 0, key_size); strcpy(key, chroms[i + j].name);
            fwrite(key, 1, key_size, f); W(f, chroms[i + j].id); W(f, chroms[i + j].size);
        }
        for (uint32_t j = c; j < block; j++) zeros(f, key_size + val_size);
        i += c;
    }
    free(key);
}

/* ---- compressed summary blocks + index (bbiWrite.c:478-536): the blocks (per_slot summaries each, independent zlib streams) are
 * compressed by all host threads, then written in order */
typedef struct { const bw_sumlist *L; uint32_t per_slot; size_t nblk, next; uint8_t **out; uLongf *out_len; } bw_sjob;
static void *bw_sworker(void *arg) {
    bw_sjob *J = (bw_sjob *)arg;
    const uLong cap = (uLong)(1.001 * (32.0 * J->per_slot) + 13);
    uint8_t *unc = (uint8_t *)malloc(32 * (size_t)J->per_slot);
    for (;;) {
        const size_t b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (b >= J->nblk) break;
        const size_t i0 = b * J->per_slot, m = J->L->n - i0 < J->per_slot ? J->L->n - i0 : J->per_slot; uint8_t *w = unc;
        for (size_t i = 0; i < m; i++) {
            const bw_sum *s = &J->L->v[i0 + i];
            memcpy(w, &s->chrom_id, 4); memcpy(w + 4, &s->start, 4); memcpy(w + 8, &s->end, 4); memcpy(w + 12, &s->valid, 4);
            memcpy(w + 16, &s->minv, 4); memcpy(w + 20, &s->maxv, 4); memcpy(w + 24, &s->sum, 4); memcpy(w + 28, &s->sumsq, 4);
            w += 32;
        }
        uint8_t *cmp = (uint8_t *)malloc(cap + 64); uLongf cl = cap;
        compress(cmp, &cl, unc, (uLong)(w - unc));
        J->out[b] = cmp; J->out_len[b] = cl;
    }
    free(unc);
    return NULL;
}
static int bw_threads(size_t njobs) {
    int T = (int)sysconf(_SC_NPROCESSORS_ONLN); if (T < 1) T = 1; if (T > 32) T = 32; if (njobs < 8) T = 1;
    return T;
}
static uint64_t write_summaries(FILE *f, bw_sumlist *L, uint32_t block, uint32_t per_slot) {
    uint32_t count = (uint32_t)L->n; W(f, count);
    bw_sjob J; J.L = L; J.per_slot = per_slot; J.nblk = (L->n + per_slot - 1) / per_slot; J.next = 0;
    J.out = (uint8_t **)calloc(J.nblk ? J.nblk : 1, sizeof(uint8_t *)); J.out_len = (uLongf *)calloc(J.nblk ? J.nblk : 1, sizeof(uLongf));
    {
        const int T = bw_threads(J.nblk); pthread_t th[32]; int started = 0;
        for (int t = 1; t < T; t++) if (pthread_create(&th[started], NULL, bw_sworker, &J) == 0) started++;
        bw_sworker(&J);
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    }
    for (size_t b = 0; b < J.nblk; b++) {
        const uint64_t pos = (uint64_t)ftello(f);
        const size_t i0 = b * per_slot, m = L->n - i0 < per_slot ? L->n - i0 : per_slot;
        for (size_t i = 0; i < m; i++) L->v[i0 + i].file_off = pos;
        fwrite(J.out[b], 1, J.out_len[b], f); free(J.out[b]);
    }
    free(J.out); free(J.out_len);
    uint64_t index_off = (uint64_t)ftello(f);
    rkey *keys = (rkey *)malloc(sizeof(rkey) * (L->n ? L->n : 1));
    for (size_t i = 0; i < L->n; i++) { keys[i].chrom = L->v[i].chrom_id; keys[i].start = L->v[i].start; keys[i].end = L->v[i].end; keys[i].off = L->v[i].file_off; }
    write_rtree(f, keys, L->n, block, per_slot, index_off);
    free(keys);
    return index_off;
}

typedef struct { bw_section *sec; size_t nsec; size_t next; uint8_t **out; uLongf *out_len; } bw_zjob;
static void *bw_zworker(void *arg) {
    bw_zjob *J = (bw_zjob *)arg;
    for (;;) {
        const size_t i0 = __atomic_fetch_add(&J->next, 16, __ATOMIC_RELAXED);
        if (i0 >= J->nsec) break;
        for (size_t i = i0; i < i0 + 16 && i < J->nsec; i++) {
            const bw_section *S = &J->sec[i]; const uint32_t n = 24 + 4u * S->count; uint8_t *buf = (uint8_t *)malloc(n);
            memcpy(buf, &S->chrom_id, 4); memcpy(buf + 4, &S->start, 4); memcpy(buf + 8, &S->end, 4); memcpy(buf + 12, &S->step, 4); memcpy(buf + 16, &S->span, 4);
            buf[20] = 3; buf[21] = 0; memcpy(buf + 22, &S->count, 2); memcpy(buf + 24, S->val, 4u * S->count);      /* type 3 = fixedStep */
            const uLong cap = (uLong)(1.001 * n + 13); uint8_t *cmp = (uint8_t *)malloc(cap + 64); uLongf cl = cap;
            compress(cmp, &cl, buf, n);
            J->out[i] = cmp; J->out_len[i] = cl; free(buf);
        }
    }
    return NULL;
}
static int sec_cmp(const void *va, const void *vb) {
    const bw_section *a = (const bw_section *)va, *b = (const bw_section *)vb;
    int d = strcmp(a->chrom, b->chrom);
    if (!d) { d = (int)a->start - (int)b->start; if (!d) d = (int)a->end - (int)b->end; }
    return d;
}
static int is_space(int c) { return c == ' ' || (c >= 9 && c <= 13); }

/* chrom_size(ctx, name) gives the chromosome size the reference would find in its size file, or -1 */
int itx_bigwig_from_wig(const char *wig_path, long (*chrom_size)(void *ctx, const char *name), void *ctx, const char *out_path, char err[ITX_ERRLEN]) {
    const uint32_t block = 256, per_slot = 1024;
    const int timing = getenv("ITX_TIMING") != NULL; struct timespec bw_t0; clock_gettime(CLOCK_MONOTONIC, &bw_t0);
#define BW_LAP(what) do { if (timing) { struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); fprintf(stderr, "[itx timing] bigWig %s: %s at %.0f ms\n", out_path, what, (t1.tv_sec - bw_t0.tv_sec) * 1e3 + (t1.tv_nsec - bw_t0.tv_nsec) / 1e6); } } while (0)
    FILE *in = fopen(wig_path, "r");
    if (!in) { snprintf(err, ITX_ERRLEN, "Couldn't open %s , %s", wig_path, strerror(errno)); return ITX_EIO; }
    /* the whole text in memory (a wiggle of `stat` is a few tens of MB: one number per consensus base), lines cut in place */
    char *text = NULL; size_t tlen = 0;
    {
        size_t cap = 1u << 20; text = (char *)malloc(cap + 1);
        for (;;) {
            if (tlen == cap) { cap *= 2; text = (char *)realloc(text, cap + 1); }
            const size_t r = fread(text + tlen, 1, cap - tlen, in);
            if (!r) break;
            tlen += r;
        }
        text[tlen] = 0;
    }
    fclose(in);
    bw_section *sec = NULL; size_t nsec = 0, csec = 0;
    char *line = NULL, *next_line = text; long lineno = 0; int rc = ITX_OK;
    char *chrom = NULL; uint32_t cs = 0, span = 0, step = 0, pos = 0; float *vals = NULL; size_t nv = 0, cv = 0;
    int in_section = 0;
    /* one fixedStep declaration's values -> sections of <= per_slot values (bwgCreate.c:226-262) */
#define FLUSH() do { \
        size_t left = nv, k = 0; uint32_t st = pos; \
        while (left) { size_t m = left > per_slot ? per_slot : left; left -= m; \
            if (nsec == csec) { csec = csec ? csec * 2 : 256; sec = (bw_section *)realloc(sec, csec * sizeof(bw_section)); } \
            bw_section *S = &sec[nsec++]; S->chrom = chrom; S->start = st; st += (uint32_t)m * step; S->end = st - step + span; S->step = step; S->span = span; \
            S->count = (uint16_t)m; S->val = (float *)malloc(sizeof(float) * m); memcpy(S->val, vals + k, sizeof(float) * m); k += m; S->chrom_id = 0; S->file_off = 0; } \
        nv = 0; } while (0)
    while (rc == ITX_OK && next_line < text + tlen) {
        line = next_line;
        lineno++;
        if (in_section && *line >= '0' && *line <= '9') {
            /* the line every wiggle of `stat` is made of: a plain count and a newline */
            const char *q = line; uint64_t acc = 0; int nd = 0;
            while (*q >= '0' && *q <= '9' && nd < 15) { acc = acc * 10 + (uint64_t)(*q - '0'); q++; nd++; }
            if (*q == '\n' && pos + (uint32_t)nv * step + span <= cs) {
                if (nv == cv) { cv = cv ? cv * 2 : 4096; vals = (float *)realloc(vals, cv * sizeof(float)); }
                vals[nv++] = (float)(double)acc;
                next_line = (char *)q + 1;
                continue;
            }
        }
        {   /* any other line: cut at its end, then word by word as before */
            char *nl = (char *)memchr(line, '\n', (size_t)(text + tlen - line));
            if (nl) { *nl = 0; next_line = nl + 1; } else next_line = text + tlen;
        }
        char *s = line; while (is_space(*s)) s++;
        if (!*s || *s == '#') continue;
        if (!in_section && nsec == 0 && (strncmp(s, "browser", 7) == 0 || strncmp(s, "track", 5) == 0)) continue;
        if (strstr(s, "chrom=")) {
            if (in_section) FLUSH();
            if (strncmp(s, "fixedStep", 9) != 0 || !is_space(s[9])) { snprintf(err, ITX_ERRLEN, "line %ld of %s: only fixedStep wiggle sections are written by iteres", lineno, wig_path); rc = ITX_EFORMAT; break; }
            uint32_t start = 0; span = 0; step = 0; chrom = NULL;
            char *q = s + 9;
            for (;;) {
                while (is_space(*q)) q++;
                if (!*q) break;
                char *w = q; while (*q && !is_space(*q)) q++;
                if (*q) *q++ = 0;
                char *eq = strchr(w, '=');
                if (!eq) { snprintf(err, ITX_ERRLEN, "strange var=val pair line %ld of %s", lineno, wig_path); rc = ITX_EFORMAT; break; }
                *eq = 0;
                if (strcmp(w, "chrom") == 0) chrom = strdup(eq + 1);
                else if (strcmp(w, "span") == 0) span = (uint32_t)strtoul(eq + 1, NULL, 10);
                else if (strcmp(w, "step") == 0) step = (uint32_t)strtoul(eq + 1, NULL, 10);
                else if (strcmp(w, "start") == 0) start = (uint32_t)strtoul(eq + 1, NULL, 10);
                else { snprintf(err, ITX_ERRLEN, "Unknown setting %s=%s line %ld of %s", w, eq + 1, lineno, wig_path); rc = ITX_EFORMAT; break; }
            }
            if (rc) break;
            if (!chrom) { snprintf(err, ITX_ERRLEN, "Missing chrom= setting line %ld of %s\n", lineno, wig_path); rc = ITX_EFORMAT; break; }
            long size = chrom_size(ctx, chrom);
            if (size < 0) { snprintf(err, ITX_ERRLEN, "hashMustFindVal: '%s' not found", chrom); rc = ITX_EFORMAT; break; }
            cs = (uint32_t)size;
            if (start > cs) { snprintf(err, ITX_ERRLEN, "line %ld of %s: chromosome %s has %u bases, but item starts at %u", lineno, wig_path, chrom, cs, start); rc = ITX_EFORMAT; break; }
            if (start == 0) { snprintf(err, ITX_ERRLEN, "Missing start= setting line %ld of %s\n", lineno, wig_path); rc = ITX_EFORMAT; break; }
            if (step == 0) { snprintf(err, ITX_ERRLEN, "Missing step= setting line %ld of %s\n", lineno, wig_path); rc = ITX_EFORMAT; break; }
            if (span == 0) span = step;
            pos = start - 1; in_section = 1; nv = 0;
            continue;
        }
        if (!in_section) { snprintf(err, ITX_ERRLEN, "Unrecognized line %ld of %s:\n%s\n", lineno, wig_path, s); rc = ITX_EFORMAT; break; }
        char *w = s; while (*s && !is_space(*s)) s++; *s = 0;
        char *endp; double v;
        {   /* wiggles of `stat` are plain counts: digits only take the short road, everything else strtod */
            const char *q = w; uint64_t acc = 0; int nd = 0;
            while (*q >= '0' && *q <= '9' && nd < 15) { acc = acc * 10 + (uint64_t)(*q - '0'); q++; nd++; }
            if (nd && !*q) { v = (double)acc; endp = (char *)q; } else v = strtod(w, &endp);
        }
        if (!*w || *endp) { snprintf(err, ITX_ERRLEN, "Expecting double field 1 line %ld of %s, got %s", lineno, wig_path, w); rc = ITX_EFORMAT; break; }
        if (pos + (uint32_t)nv * step + span > cs) {
            snprintf(err, ITX_ERRLEN, "line %ld of %s: chromosome %s has %u bases, but item ends at %u", lineno, wig_path, chrom, cs, pos + (uint32_t)nv * step + span);
            rc = ITX_EFORMAT; break;
        }
        if (nv == cv) { cv = cv ? cv * 2 : 4096; vals = (float *)realloc(vals, cv * sizeof(float)); }
        vals[nv++] = (float)v;
    }
    if (rc == ITX_OK && in_section) FLUSH();
#undef FLUSH
    BW_LAP("wig text parsed");
    free(text); free(vals);
    if (rc == ITX_OK && nsec == 0) { snprintf(err, ITX_ERRLEN, "%s is empty of data", wig_path); rc = ITX_EFORMAT; }
    if (rc != ITX_OK) { for (size_t i = 0; i < nsec; i++) free(sec[i].val); free(sec); return rc; }

    /* stable sort by (chrom, start, end) as slSort does; then the overlap check (bwgCreate.c:1062-1081) */
    {
        bw_section *tmp = (bw_section *)malloc(nsec * sizeof(bw_section));
        for (size_t w = 1; w < nsec; w *= 2) {
            for (size_t lo = 0; lo < nsec; lo += 2 * w) {
                size_t mid = lo + w < nsec ? lo + w : nsec, hi = lo + 2 * w < nsec ? lo + 2 * w : nsec, i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = sec_cmp(&sec[j], &sec[i]) < 0 ? sec[j++] : sec[i++];
                while (i < mid) tmp[k++] = sec[i++];
                while (j < hi) tmp[k++] = sec[j++];
            }
            memcpy(sec, tmp, nsec * sizeof(bw_section));
        }
        free(tmp);
    }
    for (size_t i = 0; i + 1 < nsec; i++)
        if (strcmp(sec[i].chrom, sec[i + 1].chrom) == 0 && sec[i].end > sec[i + 1].start) {
            snprintf(err, ITX_ERRLEN, "There's more than one value for %s base %d (in coordinates that start with 1).\n", sec[i].chrom, (int)sec[i + 1].start + 1);
            for (size_t k = 0; k < nsec; k++) free(sec[k].val);
            free(sec); return ITX_EFORMAT;
        }
    /* chromosome ids in order of first appearance in the sorted list */
    bw_chrom *chroms = (bw_chrom *)calloc(nsec, sizeof(bw_chrom)); uint32_t nchrom = 0, max_name = 0;
    for (size_t i = 0; i < nsec; i++) {
        if (nchrom == 0 || strcmp(sec[i].chrom, chroms[nchrom - 1].name) != 0) {
            chroms[nchrom].name = sec[i].chrom; chroms[nchrom].id = nchrom; chroms[nchrom].size = (uint32_t)chrom_size(ctx, sec[i].chrom);
            uint32_t l = (uint32_t)strlen(sec[i].chrom); if (l > max_name) max_name = l;
            nchrom++;
        }
        sec[i].chrom_id = nchrom - 1;
    }
    /* zoom levels (bwgCreate.c:824-884) */
    uint64_t total_res = 0; for (size_t i = 0; i < nsec; i++) total_res += sec[i].step;
    int min_res = (int)((total_res + nsec / 2) / nsec), initial = min_res * 10;
    uint64_t full = 0; for (size_t i = 0; i < nsec; i++) full += 24 + 4ull * sec[i].count;
    uint64_t max_reduced = full / 2, last_size = 0;
    bw_sumlist zoom[10]; memset(zoom, 0, sizeof zoom); uint32_t amount[10]; uint16_t nzoom = 0;
    for (;;) {
        reduce_sections(sec, nsec, chroms, initial, &zoom[0]);
        uint64_t size = 32ull * zoom[0].n * 2;                          /* x2: summaries compress worse than the data */
        if (size >= max_reduced && size != last_size) {
            int next = (int)(1.1 * initial * size / max_reduced);
            if (next < initial * 2) next = initial * 2;
            initial = next; last_size = size;
        } else break;
    }
    nzoom = 1; amount[0] = (uint32_t)initial;
    {
        uint64_t reduction = (uint64_t)initial;
        for (int i = 0; i < 9; i++) {
            reduction *= 4;
            if (reduction > 1000000000ull) break;
            bw_sumlist tmp; memset(&tmp, 0, sizeof tmp);
            reduce_sums(&zoom[nzoom - 1], chroms, (int)reduction, &tmp);
            uint64_t size = 32ull * tmp.n; size_t items = tmp.n;
            if (size != last_size) { zoom[nzoom] = tmp; amount[nzoom] = (uint32_t)reduction; nzoom++; } else free(tmp.v);
            if (items <= nchrom) break;
        }
    }
    BW_LAP("sections sorted, zoom levels reduced");
    FILE *f = fopen(out_path, "wb");
    if (!f) { snprintf(err, ITX_ERRLEN, "Can't open %s to write: %s", out_path, strerror(errno)); rc = ITX_EIO; goto done; }
    {
        uint32_t sig = BW_SIG, r32 = 0, unc_buf = 0; uint16_t version = BW_VERSION, r16 = 0; uint64_t r64 = 0;
        uint64_t data_off = 0, index_off = 0, ctree_off = 0, tsum_off = 0;
        W(f, sig); W(f, version); W(f, nzoom);
        long ctree_pos = ftell(f); W(f, ctree_off);
        long data_pos = ftell(f); W(f, data_off);
        long index_pos = ftell(f); W(f, index_off);
        W(f, r16); W(f, r16); W(f, r64);
        long tsum_pos = ftell(f); W(f, tsum_off);
        long unc_pos = ftell(f); W(f, unc_buf);
        W(f, r64);
        long zoom_pos[10]; uint64_t zoom_data[10], zoom_index[10];
        for (int i = 0; i < nzoom; i++) { W(f, amount[i]); W(f, r32); zoom_pos[i] = ftell(f); W(f, r64); W(f, r64); }
        tsum_off = (uint64_t)ftell(f);
        { uint64_t vc = 0; double z = 0; W(f, vc); W(f, z); W(f, z); W(f, z); W(f, z); }
        ctree_off = (uint64_t)ftell(f);
        write_chrom_tree(f, chroms, nchrom, nchrom < block ? nchrom : block, max_name);
        data_off = (uint64_t)ftell(f);
        { uint64_t sc = nsec; W(f, sc); }
        {   /* the sections are compressed by all host threads (each one is an independent zlib stream), then written in order */
            bw_zjob J; J.sec = sec; J.nsec = nsec; J.next = 0; J.out = (uint8_t **)calloc(nsec, sizeof(uint8_t *)); J.out_len = (uLongf *)calloc(nsec, sizeof(uLongf));
            int T = (int)sysconf(_SC_NPROCESSORS_ONLN); if (T < 1) T = 1; if (T > 32) T = 32; if (nsec < 64) T = 1;
            pthread_t th[32]; int started = 0;
            for (int t = 1; t < T; t++) if (pthread_create(&th[started], NULL, bw_zworker, &J) == 0) started++;
            bw_zworker(&J);
            for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
            for (size_t i = 0; i < nsec; i++) {
                sec[i].file_off = (uint64_t)ftello(f);
                fwrite(J.out[i], 1, J.out_len[i], f); free(J.out[i]);
                const uint32_t n = 24 + 4u * sec[i].count; if (n > unc_buf) unc_buf = n;
            }
            free(J.out); free(J.out_len);
        }
        BW_LAP("sections compressed and written");
        index_off = (uint64_t)ftello(f);
        {
            rkey *keys = (rkey *)malloc(sizeof(rkey) * nsec);
            for (size_t i = 0; i < nsec; i++) { keys[i].chrom = sec[i].chrom_id; keys[i].start = sec[i].start; keys[i].end = sec[i].end; keys[i].off = sec[i].file_off; }
            write_rtree(f, keys, nsec, block, 1, index_off);
            free(keys);
        }
        for (int i = 0; i < nzoom; i++) { zoom_data[i] = (uint64_t)ftello(f); zoom_index[i] = write_summaries(f, &zoom[i], block, per_slot); }
        BW_LAP("index and zoom summaries written");
        if (zoom[0].n) {
            const bw_sum *s = &zoom[0].v[0];
            uint64_t vc = s->valid; double mn = s->minv, mx = s->maxv, sd = s->sum, sq = s->sumsq;
            for (size_t i = 1; i < zoom[0].n; i++) {
                s = &zoom[0].v[i]; vc += s->valid;
                if (s->minv < mn) mn = s->minv;
                if (s->maxv > mx) mx = s->maxv;
                sd += s->sum; sq += s->sumsq;
            }
            fseek(f, (long)tsum_off, SEEK_SET);
            W(f, vc); W(f, mn); W(f, mx); W(f, sd); W(f, sq);
        } else tsum_off = 0;
        fseek(f, data_pos, SEEK_SET); W(f, data_off);
        fseek(f, index_pos, SEEK_SET); W(f, index_off);
        fseek(f, ctree_pos, SEEK_SET); W(f, ctree_off);
        fseek(f, tsum_pos, SEEK_SET); W(f, tsum_off);
        { uint32_t zmax = per_slot * 32u; if (zmax > unc_buf) unc_buf = zmax; fseek(f, unc_pos, SEEK_SET); W(f, unc_buf); }
        for (int i = 0; i < nzoom; i++) { fseek(f, zoom_pos[i], SEEK_SET); W(f, zoom_data[i]); W(f, zoom_index[i]); }
        fseek(f, 0L, SEEK_END); W(f, sig);
        if (fclose(f) != 0) { snprintf(err, ITX_ERRLEN, "write error on %s", out_path); rc = ITX_EIO; }
    }
done:
    for (int i = 0; i < 10; i++) free(zoom[i].v);
    {   /* a chromosome name is shared by its sections: free each once */
        for (uint32_t c = 0; c < nchrom; c++) free(chroms[c].name);
        for (size_t i = 0; i < nsec; i++) free(sec[i].val);
    }
    free(chroms); free(sec);
    return rc;
}

/* ---- bigWigFileCreate(inName, chromSizes, 256, 1024, FALSE, TRUE, outName) (bwgCreate.c:1088-1112): the size file is
 * read as bbiChromSizesFromFile does (bbiWrite.c:86-97): two columns per row, a later row of the same name wins */
typedef struct { char **names; long *sizes; size_t n, cap; } size_table;
static long size_lookup(void *ctx, const char *name) {
    const size_table *T = (const size_table *)ctx;
    for (size_t i = T->n; i-- > 0;) if (strcmp(T->names[i], name) == 0) return T->sizes[i];
    return -1;
}
int itx_wig_to_bigwig(const char *wig, const char *chrom_sizes, const char *bigwig, char err[ITX_ERRLEN]) {
    char lerr[ITX_ERRLEN]; if (!err) err = lerr;
    err[0] = 0;
    FILE *f = fopen(chrom_sizes, "r");
    if (!f) { snprintf(err, ITX_ERRLEN, "Couldn't open %s , %s", chrom_sizes, strerror(errno)); return ITX_EIO; }
    size_table T; memset(&T, 0, sizeof T);
    char *line = NULL; size_t cap = 0; int rc = ITX_OK; long lineno = 0;
    while (getline(&line, &cap, f) >= 0) {
        lineno++;
        char *s = line; while (is_space(*s)) s++;
        if (!*s || *s == '#') continue;
        char *name = s; while (*s && !is_space(*s)) s++;
        if (*s) *s++ = 0;
        while (is_space(*s)) s++;
        if (!*s) { snprintf(err, ITX_ERRLEN, "Expecting 2 words line %ld of %s got 1", lineno, chrom_sizes); rc = ITX_EFORMAT; break; }
        if (T.n == T.cap) { T.cap = T.cap ? T.cap * 2 : 1024; T.names = (char **)realloc(T.names, T.cap * sizeof(char *)); T.sizes = (long *)realloc(T.sizes, T.cap * sizeof(long)); }
        T.names[T.n] = strdup(name); T.sizes[T.n] = (long)strtoul(s, NULL, 10); T.n++;
    }
    free(line); fclose(f);
    if (rc == ITX_OK) rc = itx_bigwig_from_wig(wig, size_lookup, &T, bigwig, err);
    for (size_t i = 0; i < T.n; i++) free(T.names[i]);
    free(T.names); free(T.sizes);
    return rc;
}
