/* itx_host.c -- host side of libiteres_gpu: text-table loaders, dense ids, the sorted interval
 * table, BAM header -> per-tid lookup, and the output writers.  Plain C; nothing here computes
 * overlaps or counts reads (that is itx_gpu.cu).
 *
 * Reference behaviour restated here (file:line in lidaof/iteres):
 *   hashNameIntFile            cuskent/obscure.c:139-150     two-column name/int tables
 *   rmsk2binKeeperHash         generic.c:1578-1707           rmsk.txt -> elements + group tables
 *   hash iteration order       cuskent/hash.c:41-53,136-140,374-410,511-551
 *   binKeeperFirst/Next order  cuskent/binRange.c:365-392
 *   writeWigandStat/Report     generic.c:53-113
 *   writeFilterOut(/MRE)       generic.c:1709-1771,  MREwriteWigandStat generic.c:115-152
 */
#define _GNU_SOURCE
#include "itx_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <ctype.h>
#include <errno.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------ string tables */
uint32_t itx_fnv1a(const char *s, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 16777619u; }
    return h;
}
void itx_strtab_init(itx_strtab *t) {
    memset(t, 0, sizeof *t);
    t->nslot = 256; t->slot = (int32_t *)calloc(t->nslot, sizeof(int32_t));
}
void itx_strtab_free(itx_strtab *t) {
    for (int32_t i = 0; i < t->n; i++) free(t->names[i]);
    free(t->names); free(t->slot); memset(t, 0, sizeof *t);
}
int32_t itx_strtab_findn(const itx_strtab *t, const char *name, size_t len) {
    if (!t->slot) return -1;
    uint32_t m = t->nslot - 1, i = itx_fnv1a(name, len) & m;
    while (t->slot[i]) {
        int32_t k = t->slot[i] - 1;
        if (strncmp(t->names[k], name, len) == 0 && t->names[k][len] == 0) return k;
        i = (i + 1) & m;
    }
    return -1;
}
int32_t itx_strtab_find(const itx_strtab *t, const char *name) { return itx_strtab_findn(t, name, strlen(name)); }
static void strtab_place(itx_strtab *t, int32_t k) {
    const char *name = t->names[k]; size_t len = strlen(name);
    uint32_t m = t->nslot - 1, i = itx_fnv1a(name, len) & m;
    while (t->slot[i]) {
        int32_t o = t->slot[i] - 1;
        if (strcmp(t->names[o], name) == 0) break;      /* same key: the newer index takes the slot */
        i = (i + 1) & m;
    }
    t->slot[i] = k + 1;
}
int32_t itx_strtab_add(itx_strtab *t, const char *name) {
    if (t->n == t->cap) { t->cap = t->cap ? t->cap * 2 : 64; t->names = (char **)realloc(t->names, sizeof(char *) * t->cap); }
    t->names[t->n] = strdup(name);
    if ((uint32_t)(t->n + 1) * 2 > t->nslot) {
        free(t->slot); t->nslot *= 4; t->slot = (int32_t *)calloc(t->nslot, sizeof(int32_t));
        for (int32_t k = 0; k < t->n; k++) strtab_place(t, k);
    }
    strtab_place(t, t->n);
    return t->n++;
}
int32_t itx_strtab_intern(itx_strtab *t, const char *name) {
    int32_t k = itx_strtab_find(t, name);
    return k >= 0 ? k : itx_strtab_add(t, name);
}

/* Row order of the reference's tables.  A Kent hash walks buckets upward and each chain from its
 * head; new items go to the head; a doubling (when the item count exceeds the bucket count) keeps
 * the relative order inside a chain.  So the order is (hashString & (size-1)) ascending, then
 * newest first, with size = the smallest 2^k >= the item count, starting from 2^pow2_initial. */
static uint32_t kent_hash(const char *s) {
    uint32_t h = 0; int c;
    while ((c = *s++) != 0) h += (h << 3) + (uint32_t)c;       /* plain char, sign-extended like the reference */
    return h;
}
typedef struct { uint32_t bucket; int32_t seq; } ordkey;
static int ordkey_cmp(const void *a, const void *b) {
    const ordkey *x = (const ordkey *)a, *y = (const ordkey *)b;
    if (x->bucket != y->bucket) return x->bucket < y->bucket ? -1 : 1;
    return x->seq > y->seq ? -1 : (x->seq < y->seq ? 1 : 0);
}
int32_t *itx_kent_order(const itx_strtab *t, int pow2) {
    while ((int64_t)t->n > ((int64_t)1 << pow2)) pow2++;
    uint32_t mask = (uint32_t)(((uint64_t)1 << pow2) - 1);
    int32_t n = t->n > 0 ? t->n : 0;
    ordkey *k = (ordkey *)malloc(sizeof(ordkey) * (size_t)(n ? n : 1));
    for (int32_t i = 0; i < n; i++) { k[i].bucket = kent_hash(t->names[i]) & mask; k[i].seq = i; }
    qsort(k, (size_t)n, sizeof(ordkey), ordkey_cmp);
    int32_t *o = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    for (int32_t i = 0; i < n; i++) o[i] = k[i].seq;
    free(k);
    return o;
}

/* ------------------------------------------------------------------ line reader helpers */
/* split on white space in place, at most max words (chopByWhite, cuskent/common.c:1915-1953) */
static int split_white(char *s, char **w, int max) {
    int n = 0;
    while (n < max) {
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s) break;
        w[n++] = s;
        while (*s && !isspace((unsigned char)*s)) s++;
        if (!*s) break;
        *s++ = 0;
    }
    return n;
}
static int is_directory(const char *p) { struct stat st; return stat(p, &st) == 0 && S_ISDIR(st.st_mode); }

/* two-column name / integer file; lines starting with '#' and blank lines are skipped by the
 * reference's row reader; duplicates: the newest value wins on lookup */
static int load_name_int(const char *path, itx_strtab *names, int **vals, char *err) {
    FILE *f = is_directory(path) ? NULL : fopen(path, "r");
    if (!f) { snprintf(err, ITX_ERRLEN, "Couldn't open %s , %s", path, strerror(errno)); return ITX_EIO; }
    itx_strtab_init(names); *vals = NULL; int cap = 0;
    char *line = NULL; size_t lc = 0; long ln = 0; int rc = ITX_OK;
    while (getline(&line, &lc, f) >= 0) {
        ln++;
        if (line[0] == '#') continue;
        char *w[2]; int nw = split_white(line, w, 2);
        if (nw == 0) continue;
        if (nw < 2) { snprintf(err, ITX_ERRLEN, "Expecting 2 words line %ld of %s got %d", ln, path, nw); rc = ITX_EFORMAT; break; }
        if (!(w[1][0] == '-' || isdigit((unsigned char)w[1][0]))) {
            snprintf(err, ITX_ERRLEN, "Expecting number field 2 line %ld of %s, got %s", ln, path, w[1]); rc = ITX_EFORMAT; break;
        }
        int32_t k = itx_strtab_add(names, w[0]);
        if (k >= cap) { cap = cap ? cap * 2 : 64; *vals = (int *)realloc(*vals, sizeof(int) * (size_t)cap); }
        (*vals)[k] = atoi(w[1]);
    }
    free(line); fclose(f);
    return rc;
}
static int name_int_or(const itx_strtab *t, const int *vals, const char *name, int dflt) {
    int32_t k = itx_strtab_find(t, name);
    return k < 0 ? dflt : vals[k];
}

/* ------------------------------------------------------------------ rmsk loader */
typedef struct {
    int32_t chrom, start, end; uint32_t cs, ce, row; int32_t sub, fam, cla;
} raw_el;
static int raw_cmp(const void *a, const void *b) {
    const raw_el *x = (const raw_el *)a, *y = (const raw_el *)b;
    if (x->chrom != y->chrom) return x->chrom < y->chrom ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row ? 1 : 0);
}
static itx_group *grow_groups(itx_group *g, int32_t *cap, int32_t need) {
    if (need <= *cap) return g;
    int32_t nc = *cap ? *cap : 64; while (nc < need) nc *= 2;
    g = (itx_group *)realloc(g, sizeof(itx_group) * (size_t)nc);
    memset(g + *cap, 0, sizeof(itx_group) * (size_t)(nc - *cap));
    *cap = nc; return g;
}
/* finest binKeeper level (0..5) whose bin holds [s,e) entirely, or -1 (cuskent/binRange.c:119-138) */
static int bin_level(int32_t s, int32_t e, int32_t *bin) {
    int32_t a = s >> 17, b = (e - 1) >> 17;
    for (int l = 0; l < 6; l++) { if (a == b) { *bin = a; return l; } a >>= 3; b >>= 3; }
    return -1;
}

/* one piece of rmsk.txt */
typedef struct {
    char *buf; size_t lo, hi;
    const struct itx_index *ix; int filter_field; const char *filter_name, *path;
    long long row_base; long line_base;                 /* rows / lines before this piece (known after the counting pass) */
    long long n_rows; long n_lines;                     /* counting pass */
    raw_el *raw; long long nraw, rawcap, kept;          /* parsing pass: elements with LOCAL chromosome / subfamily / family / class ids */
    itx_strtab chroms, subs, fams, clas; int32_t *chrom_size; int32_t chromcap;
    itx_group *sub, *fam, *cla; int32_t subcap, famcap, clacap;
    int rc; char err[ITX_ERRLEN];
    int pass;
} rmsk_piece;
static void *rmsk_worker(void *arg) {
    rmsk_piece *P = (rmsk_piece *)arg;
    char *s = P->buf + P->lo, *end = P->buf + P->hi;
    if (P->pass == 0) {
        /* rows = lines that are neither comments nor blank (rmsk2binKeeperHash numbers them before it filters) */
        long long rows = 0; long lines = 0;
        while (s < end) {
            char *e = (char *)memchr(s, '\n', (size_t)(end - s)); if (!e) e = end;
            lines++;
            if (*s != '#') { char *q = s; while (q < e && isspace((unsigned char)*q)) q++; if (q < e) rows++; }
            s = e + 1;
        }
        P->n_rows = rows; P->n_lines = lines;
        return NULL;
    }
    const struct itx_index *ix = P->ix;
    itx_strtab_init(&P->chroms); itx_strtab_init(&P->subs); itx_strtab_init(&P->fams); itx_strtab_init(&P->clas);
    long long row = P->row_base - 1; long ln = P->line_base;
    while (s < end) {
        char *e = (char *)memchr(s, '\n', (size_t)(end - s)); if (!e) e = end;
        char *line = s; s = e + 1;
        *e = 0;                                          /* the buffer is ours; one byte past the file is allocated too */
        ln++;
        if (line[0] == '#') continue;
        char *w[17]; int nw = split_white(line, w, 17);
        if (nw == 0) continue;
        if (nw < 17) { snprintf(P->err, ITX_ERRLEN, "Expecting 17 words line %ld of %s got %d", ln, P->path, nw); P->rc = ITX_EFORMAT; return NULL; }
        row++;
        if (P->filter_field != 0 && strcmp(P->filter_name, w[P->filter_field]) != 0) continue;
        P->kept++;
        raw_el el;
        el.cs = (uint32_t)strtol(w[9][0] == '+' ? w[13] : w[15], NULL, 0);
        el.ce = (uint32_t)strtol(w[14], NULL, 0);
        el.start = (int32_t)(uint32_t)strtol(w[6], NULL, 0);
        el.end = (int32_t)(uint32_t)strtol(w[7], NULL, 0);
        el.row = (uint32_t)row;
        int32_t c = itx_strtab_find(&P->chroms, w[5]);
        if (c < 0) {
            int size = name_int_or(&ix->chromsize, ix->chromsize_val, w[5], 0);
            if (size == 0) continue;                       /* chromosome not in the size file: row dropped */
            if (size < 0) { snprintf(P->err, ITX_ERRLEN, "bad range %d,%d in binKeeperNew", 0, size); P->rc = ITX_EFORMAT; return NULL; }
            c = itx_strtab_add(&P->chroms, w[5]);
            if (c >= P->chromcap) { P->chromcap = P->chromcap ? P->chromcap * 2 : 64; P->chrom_size = (int32_t *)realloc(P->chrom_size, sizeof(int32_t) * (size_t)P->chromcap); }
            P->chrom_size[c] = size;
        }
        int32_t bin;
        if (el.start < 0 || el.end > P->chrom_size[c] || el.start > el.end) {
            snprintf(P->err, ITX_ERRLEN, "(%d %d) out of range (%d %d) in binKeeperAdd", el.start, el.end, 0, P->chrom_size[c]); P->rc = ITX_EFORMAT; return NULL;
        }
        if (bin_level(el.start, el.end, &bin) < 0) { snprintf(P->err, ITX_ERRLEN, "start %d, end %d out of range in findBin (max is 2Gb)", el.start, el.end); P->rc = ITX_EFORMAT; return NULL; }
        el.chrom = c;
        int32_t ns0 = P->subs.n, nf0 = P->fams.n;
        el.sub = itx_strtab_intern(&P->subs, w[10]);
        el.cla = itx_strtab_intern(&P->clas, w[11]);
        el.fam = itx_strtab_intern(&P->fams, w[12]);
        P->sub = grow_groups(P->sub, &P->subcap, P->subs.n);
        P->fam = grow_groups(P->fam, &P->famcap, P->fams.n);
        P->cla = grow_groups(P->cla, &P->clacap, P->clas.n);
        if (P->subs.n != ns0) { P->sub[el.sub].first_fam = el.fam; P->sub[el.sub].first_cla = el.cla; }
        if (P->fams.n != nf0) { P->fam[el.fam].first_cla = el.cla; }
        if (ix->stat_mode) {
            uint64_t len = (uint32_t)(el.end - el.start);
            P->sub[el.sub].genome_count++; P->sub[el.sub].total_length += len;
            P->fam[el.fam].genome_count++; P->fam[el.fam].total_length += len;
            P->cla[el.cla].genome_count++; P->cla[el.cla].total_length += len;
        }
        if (P->nraw == P->rawcap) { P->rawcap = P->rawcap ? P->rawcap * 2 : (1 << 14); P->raw = (raw_el *)realloc(P->raw, sizeof(raw_el) * (size_t)P->rawcap); }
        P->raw[P->nraw++] = el;
    }
    return NULL;
}
typedef struct { int fd; char *buf; size_t lo, hi; int failed; } rmsk_read_job;
static void *rmsk_read_worker(void *arg) {
    rmsk_read_job *J = (rmsk_read_job *)arg;
    size_t o = J->lo;
    while (o < J->hi) { ssize_t r = pread(J->fd, J->buf + o, J->hi - o, (off_t)o); if (r <= 0) { J->failed = 1; break; } o += (size_t)r; }
    return NULL;
}
static int rmsk_parse_parallel(struct itx_index *ix, const char *rmsk, int filter_field, const char *filter_name,
                               raw_el **raw_out, long long *nraw_out, long long *last_row, long long *kept_out, char *err) {
    int fd = is_directory(rmsk) ? -1 : open(rmsk, O_RDONLY);
    if (fd < 0) { snprintf(err, ITX_ERRLEN, "Couldn't open %s , %s", rmsk, strerror(errno)); return ITX_EIO; }
    int T = (int)sysconf(_SC_NPROCESSORS_ONLN); if (T < 1) T = 1; if (T > 64) T = 64;
    size_t n = 0; char *buf = NULL;
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
        /* a regular file: every thread reads (and so first touches) its own share of the buffer */
        n = (size_t)st.st_size; buf = (char *)malloc(n + 1);
        if (n < ((size_t)1 << 20)) T = 1;
        rmsk_read_job *J = (rmsk_read_job *)calloc((size_t)T, sizeof(rmsk_read_job)); pthread_t *rt = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
        for (int t = 0; t < T; t++) {
            J[t].fd = fd; J[t].buf = buf; J[t].lo = n / (size_t)T * (size_t)t; J[t].hi = t == T - 1 ? n : n / (size_t)T * (size_t)(t + 1);
            if (T == 1 || pthread_create(&rt[t], NULL, rmsk_read_worker, &J[t]) != 0) { rmsk_read_worker(&J[t]); rt[t] = 0; }
        }
        int bad = 0;
        for (int t = 0; t < T; t++) { if (rt[t]) pthread_join(rt[t], NULL); bad |= J[t].failed; }
        free(J); free(rt);
        if (bad) { close(fd); free(buf); snprintf(err, ITX_ERRLEN, "read error in %s", rmsk); return ITX_EIO; }
    } else {
        /* pipes and the like are read to their end */
        size_t cap = 1 << 24; buf = (char *)malloc(cap + 1);
        for (;;) { if (n == cap) { cap *= 2; buf = (char *)realloc(buf, cap + 1); } ssize_t r = read(fd, buf + n, cap - n); if (r <= 0) break; n += (size_t)r; }
        if (n < ((size_t)1 << 20)) T = 1;
    }
    close(fd);
    buf[n] = 0;
    rmsk_piece *P = (rmsk_piece *)calloc((size_t)T, sizeof(rmsk_piece));
    size_t at = 0;
    for (int t = 0; t < T; t++) {
        size_t hi = t == T - 1 ? n : n / (size_t)T * (size_t)(t + 1);
        if (hi < at) hi = at;
        while (hi < n && hi > 0 && buf[hi - 1] != '\n') hi++;             /* pieces end after a line feed */
        P[t].buf = buf; P[t].lo = at; P[t].hi = hi; P[t].ix = ix; P[t].filter_field = filter_field; P[t].filter_name = filter_name; P[t].path = rmsk;
        at = hi;
    }
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    const int timing = getenv("ITX_TIMING") != NULL; struct timespec tp0; clock_gettime(CLOCK_MONOTONIC, &tp0);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) { long long r = 0; long l = 0; for (int t = 0; t < T; t++) { P[t].row_base = r; P[t].line_base = l; r += P[t].n_rows; l += P[t].n_lines; } }
        for (int t = 0; t < T; t++) { P[t].pass = pass; if (T == 1 || pthread_create(&th[t], NULL, rmsk_worker, &P[t]) != 0) { rmsk_worker(&P[t]); th[t] = 0; } }
        for (int t = 0; t < T; t++) if (th[t]) { pthread_join(th[t], NULL); th[t] = 0; }
        if (timing) { struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); fprintf(stderr, "[itx timing] index: rmsk pass %d done after %.0f ms (%d threads)\n", pass, (t1.tv_sec - tp0.tv_sec) * 1e3 + (t1.tv_nsec - tp0.tv_nsec) * 1e-6, T); }
    }
    int rc = ITX_OK;
    for (int t = 0; t < T && rc == ITX_OK; t++) if (P[t].rc) { rc = P[t].rc; memcpy(err, P[t].err, ITX_ERRLEN); }      /* the first failure in file order */
    long long total = 0, kept = 0, rows = 0;
    for (int t = 0; t < T; t++) { total += P[t].nraw; kept += P[t].kept; rows += P[t].n_rows; }
    raw_el *raw = NULL;
    if (rc == ITX_OK) {
        raw = (raw_el *)malloc(sizeof(raw_el) * (size_t)(total ? total : 1));
        int32_t subcap = 0, famcap = 0, clacap = 0, chromcap = 0;
        long long o = 0;
        for (int t = 0; t < T; t++) {
            rmsk_piece *Q = &P[t];
            int32_t *mc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Q->chroms.n + 1)), *ms = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Q->subs.n + 1));
            int32_t *mf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Q->fams.n + 1)), *ml = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Q->clas.n + 1));
            for (int32_t k = 0; k < Q->chroms.n; k++) {
                int32_t c = itx_strtab_find(&ix->chroms, Q->chroms.names[k]);
                if (c < 0) {
                    c = itx_strtab_add(&ix->chroms, Q->chroms.names[k]);
                    if (c >= chromcap) { chromcap = chromcap ? chromcap * 2 : 64; ix->chrom_size = (int32_t *)realloc(ix->chrom_size, sizeof(int32_t) * (size_t)chromcap); }
                    ix->chrom_size[c] = Q->chrom_size[k];
                }
                mc[k] = c;
            }
            for (int32_t k = 0; k < Q->clas.n; k++) { ml[k] = itx_strtab_intern(&ix->clas, Q->clas.names[k]); ix->cla = grow_groups(ix->cla, &clacap, ix->clas.n); }
            for (int32_t k = 0; k < Q->fams.n; k++) {
                const int32_t n0 = ix->fams.n; mf[k] = itx_strtab_intern(&ix->fams, Q->fams.names[k]); ix->fam = grow_groups(ix->fam, &famcap, ix->fams.n);
                if (ix->fams.n != n0) ix->fam[mf[k]].first_cla = ml[Q->fam[k].first_cla];
            }
            for (int32_t k = 0; k < Q->subs.n; k++) {
                const int32_t n0 = ix->subs.n; ms[k] = itx_strtab_intern(&ix->subs, Q->subs.names[k]); ix->sub = grow_groups(ix->sub, &subcap, ix->subs.n);
                if (ix->subs.n != n0) { ix->sub[ms[k]].first_fam = mf[Q->sub[k].first_fam]; ix->sub[ms[k]].first_cla = ml[Q->sub[k].first_cla]; }
            }
            if (ix->stat_mode) {
                for (int32_t k = 0; k < Q->subs.n; k++) { ix->sub[ms[k]].genome_count += Q->sub[k].genome_count; ix->sub[ms[k]].total_length += Q->sub[k].total_length; }
                for (int32_t k = 0; k < Q->fams.n; k++) { ix->fam[mf[k]].genome_count += Q->fam[k].genome_count; ix->fam[mf[k]].total_length += Q->fam[k].total_length; }
                for (int32_t k = 0; k < Q->clas.n; k++) { ix->cla[ml[k]].genome_count += Q->cla[k].genome_count; ix->cla[ml[k]].total_length += Q->cla[k].total_length; }
            }
            for (long long i = 0; i < Q->nraw; i++) {
                raw_el el = Q->raw[i];
                el.chrom = mc[el.chrom]; el.sub = ms[el.sub]; el.fam = mf[el.fam]; el.cla = ml[el.cla];
                raw[o++] = el;
            }
            free(mc); free(ms); free(mf); free(ml);
        }
    }
    for (int t = 0; t < T; t++) {
        rmsk_piece *Q = &P[t];
        if (Q->pass == 1 && (Q->chroms.slot)) { itx_strtab_free(&Q->chroms); itx_strtab_free(&Q->subs); itx_strtab_free(&Q->fams); itx_strtab_free(&Q->clas); }
        free(Q->raw); free(Q->chrom_size); free(Q->sub); free(Q->fam); free(Q->cla);
    }
    free(P); free(th); free(buf);
    *raw_out = raw; *nraw_out = total; *last_row = rows - 1; *kept_out = kept;
    return rc;
}

int itx_host_index_load(struct itx_index *ix, const char *chrom_sizes, const char *rep_sizes, const char *rmsk,
                        int filter_field, const char *filter_name, char err[ITX_ERRLEN]) {
    int rc;
    ix->filter_field = filter_field; ix->stat_mode = (filter_field == 0);
    itx_strtab_init(&ix->chroms); itx_strtab_init(&ix->subs); itx_strtab_init(&ix->fams); itx_strtab_init(&ix->clas);
    itx_strtab_init(&ix->warned);
    if ((rc = load_name_int(chrom_sizes, &ix->chromsize, &ix->chromsize_val, err))) return rc;
    if ((rc = load_name_int(rep_sizes, &ix->repsize, &ix->repsize_val, err))) return rc;
    if (filter_field != 0 && (filter_field < 0 || filter_field > 16 || !filter_name)) { snprintf(err, ITX_ERRLEN, "bad filter field %d", filter_field); return ITX_EARG; }

    /* the table is read whole and parsed by all host threads: the file is cut at line ends into one piece per thread,
     * each piece is parsed into local tables (names in order of first appearance, counters, elements), and the pieces
     * are merged in file order, which reproduces the ids a sequential pass would have given */
    raw_el *raw = NULL; long long nraw = 0, row = -1, kept = 0;
    const int timing = getenv("ITX_TIMING") != NULL; struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
#define ITX_LAP(what) do { if (timing) { struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); fprintf(stderr, "[itx timing] index: %s at %.0f ms\n", what, (t1.tv_sec - ts0.tv_sec) * 1e3 + (t1.tv_nsec - ts0.tv_nsec) * 1e-6); } } while (0)
    if ((rc = rmsk_parse_parallel(ix, rmsk, filter_field, filter_name, &raw, &nraw, &row, &kept, err))) { free(raw); return rc; }
    if (filter_field != 0 && kept <= 0) {
        snprintf(err, ITX_ERRLEN, "* No repeats found related to [%s], typo? or specify wrong repName/Class/Family filter?", filter_name);
        free(raw); return ITX_EFORMAT;
    }
    ix->n_rows = row + 1; ix->n_elem = nraw; ix->n_kept = kept;
    ITX_LAP("rmsk parsed and merged");

    /* order by (chrom, start, row) unless the file already is */
    int sorted = 1;
    for (long long i = 1; i < nraw && sorted; i++) if (raw_cmp(&raw[i - 1], &raw[i]) > 0) sorted = 0;
    if (!sorted) qsort(raw, (size_t)nraw, sizeof(raw_el), raw_cmp);

    int32_t nchrom = ix->chroms.n;
    size_t ne = (size_t)(nraw ? nraw : 1);
    ix->chrom_off = (long long *)calloc((size_t)nchrom + 1, sizeof(long long));
    ix->iv = (itx_iv *)malloc(sizeof(itx_iv) * ne);
    ix->meta = (itx_meta *)malloc(sizeof(itx_meta) * ne); ix->meta2 = (itx_meta2 *)malloc(sizeof(itx_meta2) * ne);
    ix->el_chrom = (int32_t *)malloc(sizeof(int32_t) * ne);
    ix->row2el = (long long *)malloc(sizeof(long long) * (size_t)(ix->n_rows ? ix->n_rows : 1));
    for (long long r = 0; r < ix->n_rows; r++) ix->row2el[r] = -1;
    for (long long i = 0; i < nraw; i++) ix->chrom_off[raw[i].chrom + 1]++;
    for (int32_t c = 0; c < nchrom; c++) ix->chrom_off[c + 1] += ix->chrom_off[c];
    int32_t run = 0, prevc = -1;
    for (long long i = 0; i < nraw; i++) {
        const raw_el *e = &raw[i];
        if (e->chrom != prevc) { prevc = e->chrom; run = INT32_MIN; }
        if (e->end > run) run = e->end;
        ix->iv[i].start = e->start; ix->iv[i].end = e->end; ix->iv[i].pmax = run; ix->iv[i].row = e->row;
        ix->meta[i].cons_start = e->cs; ix->meta[i].cons_end = e->ce; ix->meta[i].row = e->row; ix->meta[i].sub = (uint32_t)e->sub;
        ix->meta2[i].fam = e->fam; ix->meta2[i].cla = e->cla;
        ix->el_chrom[i] = e->chrom;
        ix->row2el[e->row] = i;
    }
    free(raw);
    ITX_LAP("sorted table laid out");
    if (nraw >= 0xffffffffLL) { snprintf(err, ITX_ERRLEN, "more than 2^32 rmsk rows are not supported"); return ITX_ENOTSUP; }
    /* position buckets: first element with start >= (b << ITX_BSH), per chromosome */
    ix->chrom_bucket = (long long *)calloc((size_t)nchrom + 1, sizeof(long long));
    for (int32_t c = 0; c < nchrom; c++) ix->chrom_bucket[c + 1] = ix->chrom_bucket[c] + ((long long)ix->chrom_size[c] >> ITX_BSH) + 2;
    ix->n_bucket = ix->chrom_bucket[nchrom];
    ix->bucket = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ix->n_bucket ? ix->n_bucket : 1));
    for (int32_t c = 0; c < nchrom; c++) {
        long long i = ix->chrom_off[c], hi = ix->chrom_off[c + 1], nb = ix->chrom_bucket[c + 1] - ix->chrom_bucket[c];
        for (long long b = 0; b < nb; b++) {
            long long first = b << ITX_BSH;
            while (i < hi && (long long)ix->iv[i].start < first) i++;
            ix->bucket[ix->chrom_bucket[c] + b] = (uint32_t)i;
        }
    }

    /* per-subfamily consensus length, coverage array offsets, case-folded name classes */
    int32_t ns = ix->subs.n;
    ix->sub_len = (uint32_t *)calloc((size_t)(ns ? ns : 1), sizeof(uint32_t));
    ix->sub_bp_off = (unsigned long long *)calloc((size_t)ns + 1, sizeof(unsigned long long));
    ix->sub_fold = (int32_t *)calloc((size_t)(ns ? ns : 1), sizeof(int32_t));
    itx_strtab fold; itx_strtab_init(&fold);
    uint64_t off = 0;
    for (int32_t s = 0; s < ns; s++) {
        uint32_t L = ix->stat_mode ? (uint32_t)name_int_or(&ix->repsize, ix->repsize_val, ix->subs.names[s], 0) : 0;
        ix->sub_len[s] = L; ix->sub_bp_off[s] = off;
        if (L) off += (uint64_t)L + 1;
        char *lc2 = strdup(ix->subs.names[s]);
        for (char *p = lc2; *p; p++) *p = (char)tolower((unsigned char)*p);
        ix->sub_fold[s] = itx_strtab_intern(&fold, lc2);
        free(lc2);
    }
    ix->sub_bp_off[ns] = off; ix->bp_len = off;
    ix->sinfo = (itx_subinfo *)calloc((size_t)(ns ? ns : 1), sizeof(itx_subinfo));
    for (int32_t s = 0; s < ns; s++) { ix->sinfo[s].len = ix->sub_len[s]; ix->sinfo[s].fold = ix->sub_fold[s]; ix->sinfo[s].bp_off = ix->sub_bp_off[s]; }
    /* what the walk of an XA:Z alternate looks at (generic.c:318-334: overlap, then sameWord on the subfamily names), 16 bytes per element */
    ix->ivf = (itx_iv *)malloc(sizeof(itx_iv) * ne);
    for (long long i = 0; i < nraw; i++) { ix->ivf[i] = ix->iv[i]; ix->ivf[i].row = (uint32_t)ix->sub_fold[ix->meta[i].sub]; }
    if (ix->n_bucket >= 0xffffffffLL) { snprintf(err, ITX_ERRLEN, "genome too large for the position buckets"); return ITX_ENOTSUP; }
    ix->cinfo = (itx_chrominfo *)calloc((size_t)(nchrom ? nchrom : 1), sizeof(itx_chrominfo));
    for (int32_t c = 0; c < nchrom; c++) {
        ix->cinfo[c].off = (uint32_t)ix->chrom_off[c]; ix->cinfo[c].bucket = (uint32_t)ix->chrom_bucket[c];
        ix->cinfo[c].size = ix->chrom_size[c]; ix->cinfo[c].n = (uint32_t)(ix->chrom_off[c + 1] - ix->chrom_off[c]);
    }
    itx_strtab_free(&fold);
    ITX_LAP("buckets and subfamily tables done");
    return ITX_OK;
}

uint32_t *itx_names32(char *const *names, int32_t n) {
    uint32_t *t = (uint32_t *)calloc((size_t)(n > 0 ? n : 1) * 8, 4);
    for (int32_t c = 0; c < n; c++) {
        const size_t l = strlen(names[c]);
        if (l >= 32) memset(t + 8 * (size_t)c, 0xff, 32);
        else memcpy(t + 8 * (size_t)c, names[c], l);            /* (little-endian host, like the device) */
    }
    return t;
}

void itx_host_index_free(struct itx_index *ix) {
    itx_strtab_free(&ix->chromsize); free(ix->chromsize_val);
    itx_strtab_free(&ix->repsize); free(ix->repsize_val);
    itx_strtab_free(&ix->chroms); free(ix->chrom_size); free(ix->chrom_off);
    if (ix->el_names) {
        for (long long i = 0; i < ix->n_elem; i++) { for (uint32_t k = 0; k < ix->el_names_n[i]; k++) free(ix->el_names[i][k]); free(ix->el_names[i]); }
        free(ix->el_names); free(ix->el_names_n); free(ix->el_names_cap);
    }
    itx_strtab_free(&ix->subs); itx_strtab_free(&ix->fams); itx_strtab_free(&ix->clas); itx_strtab_free(&ix->warned);
    free(ix->sub); free(ix->fam); free(ix->cla); free(ix->sub_len); free(ix->sub_bp_off); free(ix->sub_fold);
    free(ix->iv); free(ix->ivf); free(ix->bucket); free(ix->chrom_bucket); free(ix->cinfo); free(ix->sinfo); free(ix->meta); free(ix->meta2); free(ix->el_chrom); free(ix->row2el);
    free(ix->bp); free(ix->bp_u); free(ix->bp_cpg); free(ix->el_cnt); free(ix->el_cnt_u); free(ix->el_cpg); free(ix->el_cpg_score);
    free(ix->row_cnt); free(ix->row_cnt_u);
    free(ix->sub_order); free(ix->fam_order); free(ix->cla_order);
}

/* ------------------------------------------------------------------ BAM header (bam_header_read, cussamtools/bam.c:69-109) */
int itx_host_parse_bam_header(struct itx_index *ix, const uint8_t *bam, uint64_t len, int addChr,
                              struct itx_bam_header *h, char err[ITX_ERRLEN]) {
    memset(h, 0, sizeof *h);
    if (len < 12 || memcmp(bam, "BAM\1", 4) != 0) { snprintf(err, ITX_ERRLEN, "invalid BAM binary header (this is not a BAM file)"); return ITX_EFORMAT; }
    int32_t l_text; memcpy(&l_text, bam + 4, 4);
    uint64_t p = 8 + (uint64_t)(uint32_t)l_text;
    if (l_text < 0 || p + 4 > len) { snprintf(err, ITX_ERRLEN, "truncated BAM header"); return ITX_EFORMAT; }
    int32_t n_ref; memcpy(&n_ref, bam + p, 4); p += 4;
    if (n_ref < 0) { snprintf(err, ITX_ERRLEN, "truncated BAM header"); return ITX_EFORMAT; }
    h->n_ref = n_ref; h->addChr = addChr;
    h->names = (char **)calloc((size_t)(n_ref ? n_ref : 1), sizeof(char *));
    h->lens = (uint32_t *)calloc((size_t)(n_ref ? n_ref : 1), sizeof(uint32_t));
    h->tid = (itx_tidinfo *)calloc((size_t)(n_ref ? n_ref : 1), sizeof(itx_tidinfo));
    for (int32_t i = 0; i < n_ref; i++) {
        if (p + 4 > len) goto trunc;
        int32_t l; memcpy(&l, bam + p, 4); p += 4;
        if (l < 0 || p + (uint64_t)l + 4 > len) goto trunc;
        h->names[i] = (char *)malloc((size_t)l + 1); memcpy(h->names[i], bam + p, (size_t)l); h->names[i][l] = 0; p += (uint64_t)l;
        memcpy(&h->lens[i], bam + p, 4); p += 4;
        /* the chromosome name the reference would look up (generic.c:781-791) */
        const char *nm = h->names[i]; char buf[512];
        itx_tidinfo *t = &h->tid[i]; t->flags = 0; t->chrom = -1; t->cend = 0; t->csid = -1;
        if (addChr) {
            if (strncmp(nm, "GL", 2) == 0) { t->flags |= ITX_TID_GLSKIP; continue; }
            else if (strcasecmp(nm, "MT") == 0) { snprintf(buf, sizeof buf, "chrM"); nm = buf; }
            else if (strncmp(nm, "chr", 3) != 0) { snprintf(buf, sizeof buf, "chr%s", nm); nm = buf; }
        }
        int size = name_int_or(&ix->chromsize, ix->chromsize_val, nm, 2);
        uint32_t cend = (uint32_t)(size - 1);
        if (cend == 1) { t->flags |= ITX_TID_UNKNOWN; continue; }
        t->cend = cend;
        t->chrom = itx_strtab_find(&ix->chroms, nm);
        t->csid = itx_strtab_find(&ix->chromsize, nm);
    }
    h->hdr_len = p;
    return ITX_OK;
trunc:
    snprintf(err, ITX_ERRLEN, "truncated BAM header");
    return ITX_EFORMAT;
}

/* ------------------------------------------------------------------ writers */
static double rpkm(uint64_t reads, uint64_t total_length, uint64_t mapped) { return reads / (mapped * 1e-9 * total_length); }
static double rpm(uint64_t reads, uint64_t mapped) { return reads / (mapped * 1e-6); }

static void ensure_orders(struct itx_index *ix) {
    if (!ix->sub_order) ix->sub_order = itx_kent_order(&ix->subs, 12);
    if (!ix->fam_order) ix->fam_order = itx_kent_order(&ix->fams, 12);
    if (!ix->cla_order) ix->cla_order = itx_kent_order(&ix->clas, 12);
}
#define LLU(x) ((unsigned long long)(x))

/* n lines "%u\n": millions of them per wiggle file, so the digits are laid down by hand into a block buffer */
static void put_u32_lines(FILE *f, const uint32_t *v, uint32_t n) {
    char buf[1 << 16]; size_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (w + 12 > sizeof buf) { fwrite(buf, 1, w, f); w = 0; }
        uint32_t x = v[i]; char t[10]; int k = 0;
        do { t[k++] = (char)('0' + x % 10u); x /= 10u; } while (x);
        while (k) buf[w++] = t[--k];
        buf[w++] = '\n';
    }
    if (w) fwrite(buf, 1, w, f);
}

int itx_write_stat(itx_index *ix, const char *subfam_stat, const char *wig, const char *fam_stat,
                   const char *class_stat, const char *wig_unique, uint64_t N, uint64_t NU) {
    if (!ix->stat_mode) return ITX_EARG;
    ensure_orders(ix);
    FILE *fs = fopen(subfam_stat, "w"), *fw = fopen(wig, "w"), *fu = fopen(wig_unique, "w");
    if (!fs || !fw || !fu) { if (fs) fclose(fs); if (fw) fclose(fw); if (fu) fclose(fu); return ITX_EIO; }
    static const char *RATE_COLS = "all_reads_RPKM\tall_reads_RPM\tunique_reads_RPKM\tunique_reads_RPM";
    fprintf(fs, "#subfamily\tfamily\tclass\tconsensus_length\treads_count\tunique_reads_count\ttotal_length\tgenome_count\t%s\n", RATE_COLS);
    for (int32_t k = 0; k < ix->subs.n; k++) {
        int32_t s = ix->sub_order[k]; const itx_group *g = &ix->sub[s];
        fprintf(fs, "%s\t%s\t%s\t%u\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->subs.names[s],
                ix->fams.names[g->first_fam], ix->clas.names[g->first_cla], ix->sub_len[s], LLU(g->read_count),
                LLU(g->read_count_unique), LLU(g->total_length), LLU(g->genome_count), rpkm(g->read_count, g->total_length, N),
                rpm(g->read_count, N), rpkm(g->read_count_unique, g->total_length, NU), rpm(g->read_count_unique, NU));
        uint32_t L = ix->sub_len[s];
        if (L) {
            fprintf(fw, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[s]);
            fprintf(fu, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[s]);
            const uint32_t *a = ix->bp + ix->sub_bp_off[s], *u = ix->bp_u + ix->sub_bp_off[s];
            put_u32_lines(fw, a, L); put_u32_lines(fu, u, L);           /* "%u\n" per consensus base */
        }
    }
    fclose(fw); fclose(fs); fclose(fu);
    FILE *ff = fopen(fam_stat, "w"); if (!ff) return ITX_EIO;
    fprintf(ff, "#family\tclass\treads_count\tunique_reads_count\ttotal_length\tgenome_count\t%s\n", RATE_COLS);
    for (int32_t k = 0; k < ix->fams.n; k++) {
        int32_t s = ix->fam_order[k]; const itx_group *g = &ix->fam[s];
        fprintf(ff, "%s\t%s\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->fams.names[s], ix->clas.names[g->first_cla],
                LLU(g->read_count), LLU(g->read_count_unique), LLU(g->total_length), LLU(g->genome_count),
                rpkm(g->read_count, g->total_length, N), rpm(g->read_count, N), rpkm(g->read_count_unique, g->total_length, NU), rpm(g->read_count_unique, NU));
    }
    fclose(ff);
    FILE *fc = fopen(class_stat, "w"); if (!fc) return ITX_EIO;
    fprintf(fc, "#class\treads_count\tunique_reads_count\ttotal_length\tgenome_count\t%s\n", RATE_COLS);
    for (int32_t k = 0; k < ix->clas.n; k++) {
        int32_t s = ix->cla_order[k]; const itx_group *g = &ix->cla[s];
        fprintf(fc, "%s\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->clas.names[s],
                LLU(g->read_count), LLU(g->read_count_unique), LLU(g->total_length), LLU(g->genome_count),
                rpkm(g->read_count, g->total_length, N), rpm(g->read_count, N), rpkm(g->read_count_unique, g->total_length, NU), rpm(g->read_count_unique, NU));
    }
    fclose(fc);
    return ITX_OK;
}

int itx_write_report(const char *path, const uint64_t cnt[13], uint32_t mapQ, const char *subfam) {
    FILE *f = fopen(path, "w"); if (!f) return ITX_EIO;
    fprintf(f, "total reads (pair): %llu\n", LLU(cnt[0]));
    fprintf(f, "mappable reads (pair): %llu\n", LLU(cnt[6]));
    fprintf(f, "uniquely mapped reads (pair) (mapQ >= %u): %llu\n", mapQ, LLU(cnt[7]));
    fprintf(f, "non-redundant uniquely mapped reads (pair): %llu\n", LLU(cnt[11]));
    fprintf(f, "mapped reads (pair) overlap with repeats but discarded due to mapped to different subfamilies: %llu\n", LLU(cnt[12]));
    fprintf(f, "mapped reads (pair) overlap with [%s] repeats: %llu\n", subfam, LLU(cnt[9]));
    fprintf(f, "uniquely mapped reads (pair) overlap with [%s] repeats: %llu\n", subfam, LLU(cnt[10]));
    fclose(f);
    return ITX_OK;
}

/* Locus order of the reference's per-element tables: chromosomes in hash order, then binKeeper bins
 * by global bin id (coarse levels first: offsets 0,1,9,73,585,4681 for levels 5..0), newest row
 * first inside a bin. */
typedef struct { int32_t chrom_rank; int32_t bin; uint32_t row; long long el; } lockey;
static int lockey_cmp(const void *a, const void *b) {
    const lockey *x = (const lockey *)a, *y = (const lockey *)b;
    if (x->chrom_rank != y->chrom_rank) return x->chrom_rank < y->chrom_rank ? -1 : 1;
    if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
    return x->row > y->row ? -1 : (x->row < y->row ? 1 : 0);
}
static lockey *locus_order(struct itx_index *ix) {
    static const int32_t LEVEL_OFFSET[6] = {4681, 585, 73, 9, 1, 0};
    int32_t *co = itx_kent_order(&ix->chroms, 12);
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ix->chroms.n ? ix->chroms.n : 1));
    for (int32_t k = 0; k < ix->chroms.n; k++) rank[co[k]] = k;
    lockey *K = (lockey *)malloc(sizeof(lockey) * (size_t)(ix->n_elem ? ix->n_elem : 1));
    for (long long i = 0; i < ix->n_elem; i++) {
        int32_t b = 0; int l = bin_level(ix->iv[i].start, ix->iv[i].end, &b);
        if (l < 0) l = 5;
        K[i].chrom_rank = rank[ix->el_chrom[i]]; K[i].bin = LEVEL_OFFSET[l] + b; K[i].row = ix->meta[i].row; K[i].el = i;
    }
    qsort(K, (size_t)ix->n_elem, sizeof(lockey), lockey_cmp);
    free(co); free(rank);
    return K;
}

int itx_write_filter(itx_index *ix, const char *path, int readlist, int threshold, uint64_t N) {
    if (!ix->el_cnt) return ITX_EARG;
    if (readlist && !ix->el_names) return ITX_ENOTSUP;
    FILE *f = fopen(path, "w"); if (!f) return ITX_EIO;
    fprintf(f, "#chr\tstart\tend\tlength\trepName\trepClass\trepFamily\treadsCount\tRPKM\tRPM%s\n", readlist ? "\treadsList" : "");
    lockey *K = locus_order(ix);
    for (long long k = 0; k < ix->n_elem; k++) {
        long long i = K[k].el; int count = (int)ix->el_cnt[i];
        if (count < threshold) continue;
        int32_t s = ix->iv[i].start, e = ix->iv[i].end; uint32_t len = (uint32_t)(e - s);
        fprintf(f, "%s\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%.3f\t%.3f", ix->chroms.names[ix->el_chrom[i]], s, e, (int)len,
                ix->subs.names[ix->meta[i].sub], ix->clas.names[ix->meta2[i].cla], ix->fams.names[ix->meta2[i].fam], count,
                rpkm((uint64_t)(unsigned long long)count, (uint64_t)len, N), rpm((uint64_t)(unsigned long long)count, N));
        if (readlist) {
            fputc('\t', f);
            for (uint32_t r = 0; r < ix->el_names_n[i]; r++) { if (r) fputc(',', f); fputs(ix->el_names[i][r], f); }
        }
        fputc('\n', f);
    }
    free(K); fclose(f);
    return ITX_OK;
}

int itx_write_cpg_stat(itx_index *ix, const char *subfam_stat, const char *wig, const char *fam_stat, const char *class_stat) {
    if (!ix->stat_mode) return ITX_EARG;
    ensure_orders(ix);
    FILE *fs = fopen(subfam_stat, "w"), *fw = fopen(wig, "w");
    if (!fs || !fw) { if (fs) fclose(fs); if (fw) fclose(fw); return ITX_EIO; }
    fprintf(fs, "#subfamily\tfamily\tclass\tconsensus_length\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->subs.n; k++) {
        int32_t s = ix->sub_order[k]; const itx_group *g = &ix->sub[s]; uint32_t L = ix->sub_len[s];
        fprintf(fs, "%s\t%s\t%s\t%u\t%u\t%.4f\t%llu\t%llu\n", ix->subs.names[s], ix->fams.names[g->first_fam], ix->clas.names[g->first_cla],
                L, g->cpg_count, g->cpg_score, LLU(g->total_length), LLU(g->genome_count));
        if (L) {
            fprintf(fw, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[s]);
            const double *v = ix->bp_cpg + ix->sub_bp_off[s];
            for (uint32_t m = 0; m < L; m++) fprintf(fw, "%.4f\n", v[m]);
        }
    }
    fclose(fw); fclose(fs);
    FILE *ff = fopen(fam_stat, "w"); if (!ff) return ITX_EIO;
    fprintf(ff, "#family\tclass\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->fams.n; k++) {
        int32_t s = ix->fam_order[k]; const itx_group *g = &ix->fam[s];
        fprintf(ff, "%s\t%s\t%u\t%.4f\t%llu\t%llu\n", ix->fams.names[s], ix->clas.names[g->first_cla], g->cpg_count, g->cpg_score, LLU(g->total_length), LLU(g->genome_count));
    }
    fclose(ff);
    FILE *fc = fopen(class_stat, "w"); if (!fc) return ITX_EIO;
    fprintf(fc, "#class\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->clas.n; k++) {
        int32_t s = ix->cla_order[k]; const itx_group *g = &ix->cla[s];
        fprintf(fc, "%s\t%u\t%.4f\t%llu\t%llu\n", ix->clas.names[s], g->cpg_count, g->cpg_score, LLU(g->total_length), LLU(g->genome_count));
    }
    fclose(fc);
    return ITX_OK;
}

int itx_write_cpg_filter(itx_index *ix, const char *path, double thr) {
    if (!ix->el_cpg) return ITX_EARG;
    FILE *f = fopen(path, "w"); if (!f) return ITX_EIO;
    fprintf(f, "#chr\tstart\tend\tlength\trepName\trepClass\trepFamily\tcovered_CpG_site\ttotal_CpG_score\n");
    lockey *K = locus_order(ix);
    for (long long k = 0; k < ix->n_elem; k++) {
        long long i = K[k].el;
        if (!(ix->el_cpg_score[i] > thr)) continue;
        int32_t s = ix->iv[i].start, e = ix->iv[i].end;
        fprintf(f, "%s\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%.3f\n", ix->chroms.names[ix->el_chrom[i]], s, e, (int)(uint32_t)(e - s),
                ix->subs.names[ix->meta[i].sub], ix->clas.names[ix->meta2[i].cla], ix->fams.names[ix->meta2[i].fam],
                (int)ix->el_cpg[i], ix->el_cpg_score[i]);
    }
    free(K); fclose(f);
    return ITX_OK;
}

/* ------------------------------------------------------------------ one BAM across ranks: the chain between the parts */
/* Rank 0's chain starts behind the BAM header, so its report stands.  Every later rank guessed (or was told) where its first
 * record starts; it is right when that is where the chain of the ranks before it arrives.  A part the chain runs over entirely
 * (a record longer than the part) must have met no record start; a part without blocks is transparent; once the chain has ended
 * (a cut record) nothing after it counts, as in the reference, whose loop stops there (generic.c:745).
 * Returns the first rank whose entry is wrong and, in *forced_entry, the entry it has to scan again from -- or -1. */
int itx_shard_chain_check(int nranks, const itx_shard_report *rep, uint64_t *forced_entry) {
    if (nranks < 2) return -1;
    uint64_t x = rep[0].exit_rel;                    /* where the chain stands, relative to the first own byte of the next rank with blocks */
    for (int k = 1; k < nranks; k++) {
        if (rep[k].own_bytes == 0) continue;
        if (x == ITX_OFF_NONE) { x = rep[k].exit_rel; continue; }             /* nothing is known (no rank before this one had a record): its guess stands */
        if (x == ITX_OFF_END) {
            if (rep[k].entry_rel != ITX_OFF_END) { *forced_entry = ITX_OFF_END; return k; }
            continue;
        }
        if (x >= rep[k].own_bytes) {                 /* the chain runs over the whole part */
            if (rep[k].entry_rel != ITX_OFF_NONE && rep[k].entry_rel != x) { *forced_entry = x; return k; }
            x -= rep[k].own_bytes;
            continue;
        }
        if (rep[k].entry_rel != x) { *forced_entry = x; return k; }
        x = rep[k].exit_rel;
    }
    return -1;
}

/* ------------------------------------------------------------------ accessors */
int32_t itx_n_subfam(const itx_index *ix) { return ix->stat_mode ? ix->subs.n : 0; }
int32_t itx_n_fam(const itx_index *ix) { return ix->stat_mode ? ix->fams.n : 0; }
int32_t itx_n_class(const itx_index *ix) { return ix->stat_mode ? ix->clas.n : 0; }
int64_t itx_n_elem(const itx_index *ix) { return ix->n_elem; }
int64_t itx_n_rows(const itx_index *ix) { return ix->n_rows; }
int64_t itx_n_repeats_parsed(const itx_index *ix) { return ix->n_kept; }
int32_t itx_n_chrom(const itx_index *ix) { return ix->chroms.n; }
static int32_t ordered(const itx_index *cix, int which, int32_t i) {
    struct itx_index *ix = (struct itx_index *)cix; ensure_orders(ix);
    return which == 0 ? ix->sub_order[i] : which == 1 ? ix->fam_order[i] : ix->cla_order[i];
}
const char *itx_name(const itx_index *ix, int which, int32_t i) {
    int32_t k = ordered(ix, which, i);
    return which == 0 ? ix->subs.names[k] : which == 1 ? ix->fams.names[k] : ix->clas.names[k];
}
void itx_counts(const itx_index *ix, int which, int32_t i, uint64_t out[4]) {
    int32_t k = ordered(ix, which, i);
    const itx_group *g = which == 0 ? &ix->sub[k] : which == 1 ? &ix->fam[k] : &ix->cla[k];
    out[0] = g->read_count; out[1] = g->read_count_unique; out[2] = g->total_length; out[3] = g->genome_count;
}
uint32_t itx_subfam_length(const itx_index *ix, int32_t i) { return ix->sub_len[ordered(ix, 0, i)]; }
const uint32_t *itx_subfam_bp(const itx_index *ix, int32_t i, int unique) {
    int32_t k = ordered(ix, 0, i);
    if (!ix->bp) return NULL;
    return (unique ? ix->bp_u : ix->bp) + ix->sub_bp_off[k];
}
const uint32_t *itx_elem_counts_by_row(itx_index *ix, int unique) {
    if (!ix->el_cnt) return NULL;
    uint32_t **dst = unique ? &ix->row_cnt_u : &ix->row_cnt;
    free(*dst); *dst = (uint32_t *)calloc((size_t)(ix->n_rows ? ix->n_rows : 1), sizeof(uint32_t));
    const uint32_t *src = unique ? ix->el_cnt_u : ix->el_cnt;
    for (long long i = 0; i < ix->n_elem; i++) (*dst)[ix->meta[i].row] = src[i];
    return *dst;
}
void itx_scan_opts_default(itx_scan_opts *o) {
    memset(o, 0, sizeof *o);
    o->mapQ = 10; o->iSize = 500; o->extension = 150; o->minCoverage = 1e-4f; o->diffSubfam = 1;
}
const char *itx_version(void) { return "iteres-b200 0.1 (path of iteres 0.3.3-r123)"; }
