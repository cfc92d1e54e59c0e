/* iteres_oracle.c -- plain-C restatement of the iteres hot path.  TEST INFRASTRUCTURE ONLY
 * (see iteres_oracle.h for who may load it and how its parity is pinned).
 *
 * It deliberately keeps the reference's data structures in spirit -- Kent's binKeeper with LIFO
 * bin lists and a level-by-level scan, chained string hashes whose iteration order defines the
 * output row order -- so that it checks the GPU path's different layout (start-sorted arrays and
 * an explicit order key) rather than sharing its assumptions.  No line is copied: the reference's
 * behaviour is restated from SURVEY.md Appendix A-C and the cited lines.
 */
#define _GNU_SOURCE
#include "iteres_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <ctype.h>
#include <errno.h>
#include <zlib.h>

/* ------------------------------------------------------------------------------------------
 * Kent string hash, restated (cuskent/hash.c:41-53 hashString; 115-142 hashAddN incl. doubling;
 * 374-410 hashResize keeps in-bucket order; 511-551 iteration = bucket ascending, newest first).
 * Names are kept in insertion order; lookups go through a private open-addressing map; the
 * Kent iteration order is materialised on demand.
 * ------------------------------------------------------------------------------------------ */
static uint32_t kent_hash_string(const char *s) {
    uint32_t h = 0; int c;
    while ((c = *s++) != '\0') h += (h << 3) + (uint32_t)c;     /* c is a (signed) char promoted to int */
    return h;
}

typedef struct {
    char **names; int32_t n, cap;
    int32_t *slot; uint32_t nslot;        /* open addressing: index+1, 0 = empty */
    int pow2;                              /* Kent table size exponent (starts at 12 or the given one) */
} ktab;

static void ktab_init(ktab *t, int pow2) { memset(t, 0, sizeof *t); t->pow2 = pow2 ? pow2 : 12; t->nslot = 1024; t->slot = calloc(t->nslot, sizeof(int32_t)); }
static void ktab_free(ktab *t) { for (int32_t i = 0; i < t->n; i++) free(t->names[i]); free(t->names); free(t->slot); memset(t, 0, sizeof *t); }
static uint32_t fnv(const char *s) { uint32_t h = 2166136261u; while (*s) { h ^= (uint8_t)*s++; h *= 16777619u; } return h; }
static int32_t ktab_find(const ktab *t, const char *name) {
    uint32_t m = t->nslot - 1, i = fnv(name) & m;
    while (t->slot[i]) { int32_t k = t->slot[i] - 1; if (strcmp(t->names[k], name) == 0) return k; i = (i + 1) & m; }
    return -1;
}
static void ktab_rehash(ktab *t) {
    free(t->slot); t->nslot *= 4; t->slot = calloc(t->nslot, sizeof(int32_t));
    uint32_t m = t->nslot - 1;
    for (int32_t k = 0; k < t->n; k++) { uint32_t i = fnv(t->names[k]) & m; while (t->slot[i]) i = (i + 1) & m; t->slot[i] = k + 1; }
}
/* append without a duplicate check (hashAdd semantics); the private map keeps the NEWEST index,
 * which is what hashLookup would return (chain head). */
static int32_t ktab_add(ktab *t, const char *name) {
    if (t->n == t->cap) { t->cap = t->cap ? t->cap * 2 : 64; t->names = realloc(t->names, sizeof(char *) * t->cap); }
    t->names[t->n] = strdup(name);
    if ((uint32_t)(t->n + 1) * 2 > t->nslot) ktab_rehash(t);
    uint32_t m = t->nslot - 1, i = fnv(name) & m;
    while (t->slot[i]) { int32_t k = t->slot[i] - 1; if (strcmp(t->names[k], name) == 0) break; i = (i + 1) & m; }
    t->slot[i] = t->n + 1;
    t->n++;
    /* Kent doubles when elCount > size (expansionFactor 1.0) */
    while (t->n > (1 << t->pow2)) t->pow2++;
    return t->n - 1;
}
typedef struct { uint32_t bucket; int32_t seq; } kord;
static int kord_cmp(const void *a, const void *b) {
    const kord *x = a, *y = b;
    if (x->bucket != y->bucket) return x->bucket < y->bucket ? -1 : 1;
    return y->seq - x->seq;                       /* newest first inside a bucket */
}
/* order[] gets the insertion indices in hashFirst/hashNext order */
static int32_t *ktab_order(const ktab *t) {
    kord *k = malloc(sizeof(kord) * (t->n ? t->n : 1));
    uint32_t mask = (1u << t->pow2) - 1;
    for (int32_t i = 0; i < t->n; i++) { k[i].bucket = kent_hash_string(t->names[i]) & mask; k[i].seq = i; }
    qsort(k, t->n, sizeof(kord), kord_cmp);
    int32_t *o = malloc(sizeof(int32_t) * (t->n ? t->n : 1));
    for (int32_t i = 0; i < t->n; i++) o[i] = k[i].seq;
    free(k);
    return o;
}

/* ------------------------------------------------------------------------------------------
 * binKeeper, restated (cuskent/binRange.c:20-25 offsets/shifts, 119-138 bin choice, 140-155 new,
 * 171-186 add (LIFO), 196-227 find (result built by head insertion), 365-392 first/next).
 * ------------------------------------------------------------------------------------------ */
static const int BIN_OFFSETS[6] = {4096 + 512 + 64 + 8 + 1, 512 + 64 + 8 + 1, 64 + 8 + 1, 8 + 1, 1, 0};
#define BIN_FIRST_SHIFT 17
#define BIN_NEXT_SHIFT 3

static int bin_from_range(int start, int end) {
    int sb = start >> BIN_FIRST_SHIFT, eb = (end - 1) >> BIN_FIRST_SHIFT;
    for (int i = 0; i < 6; i++) {
        if (sb == eb) return BIN_OFFSETS[i] + sb;
        sb >>= BIN_NEXT_SHIFT; eb >>= BIN_NEXT_SHIFT;
    }
    return -1;
}

typedef struct {
    int start, end;                        /* genomic */
    uint32_t cons_start, cons_end, length;
    int32_t sub, fam, cla;                 /* ids into the three name tables, by the ROW's own strings */
    int32_t chrom;                         /* id in the rmsk chromosome table */
    int32_t row;                           /* 0-based index among parsed rmsk rows */
    int64_t next;                          /* next element in the same bin (older) */
    /* filter mode: read names in arrival order (the reference prepends and reverses at print time) */
    uint32_t nreads, nreads_unique;
    char **readnames; uint32_t readcap;
    uint32_t cpgCount; double cpgTotalScore;
    char *name, *fname, *cname;            /* the row's own strings */
} oelem;

typedef struct { int minPos, maxPos, binCount; int64_t *binHead; } obk;

typedef struct {
    uint64_t read_count, read_count_unique, genome_count, total_length;
    uint32_t length; uint32_t *bp_total, *bp_total_unique; double *cpgScore;
    uint32_t cpgCount; double cpgTotalScore;
    int32_t fam_first, cla_first;          /* family/class strings of the FIRST row with this name (generic.c:1638-1640) */
    char *fname, *cname;
} osub;
typedef struct { uint64_t read_count, read_count_unique, genome_count, total_length; uint32_t cpgCount; double cpgTotalScore; char *cname; } ofam;

struct ora_index {
    ktab chromsize_names; int *chromsize_val;        /* hashNameIntFile(chrom sizes) */
    ktab repsize_names; int *repsize_val;            /* hashNameIntFile(repeat sizes) */
    ktab chroms; obk *bks; int32_t bk_cap;           /* hashRmsk: chromosome -> binKeeper */
    ktab subs; osub *sub; int32_t sub_cap;
    ktab fams; ofam *fam; int32_t fam_cap;
    ktab clas; ofam *cla; int32_t cla_cap;
    oelem *el; int64_t n_el, el_cap;
    int filter_field;
    /* state that persists across files of one run */
    ktab nochr; ktab dup; char dupkey[256];
    uint64_t cnt[13];
    int32_t *sub_order, *fam_order, *cla_order;      /* cached Kent iteration orders */
};

/* chopByWhite (cuskent/common.c:1915-1953): split in place, at most max words */
static int chop_white(char *in, char **out, int max) {
    int n = 0;
    for (;;) {
        if (n >= max) break;
        while (isspace((unsigned char)*in)) ++in;
        if (*in == 0) break;
        out[n++] = in;
        while (*in && !isspace((unsigned char)*in)) ++in;
        if (*in == 0) break;
        *in++ = 0;
    }
    return n;
}

/* hashNameIntFile (obscure.c:139-150): two-column file; duplicates allowed, newest wins on lookup */
static int load_name_int(const char *path, ktab *names, int **vals, char err[256]) {
    FILE *f = fopen(path, "r");
    if (!f) { if (err) snprintf(err, 256, "Couldn't open %s , %s", path, strerror(errno)); return -1; }
    ktab_init(names, 16);
    int cap = 0; *vals = NULL;
    char *line = NULL; size_t lc = 0; ssize_t n; long lineno = 0;
    while ((n = getline(&line, &lc, f)) >= 0) {
        lineno++;
        if (line[0] == '#') continue;
        char *w[2]; int nw = chop_white(line, w, 2);
        if (nw == 0) continue;
        if (nw < 2) { if (err) snprintf(err, 256, "Expecting 2 words line %ld of %s got %d", lineno, path, nw); free(line); fclose(f); return -1; }
        if (w[1][0] != '-' && !isdigit((unsigned char)w[1][0])) { if (err) snprintf(err, 256, "Expecting number field 2 line %ld of %s, got %s", lineno, path, w[1]); free(line); fclose(f); return -1; }
        int32_t k = ktab_add(names, w[0]);
        if (k >= cap) { cap = cap ? cap * 2 : 64; *vals = realloc(*vals, sizeof(int) * cap); }
        (*vals)[k] = atoi(w[1]);
    }
    free(line); fclose(f);
    return 0;
}
static int name_int_default(const ktab *t, const int *vals, const char *name, int dflt) {
    int32_t k = ktab_find(t, name); return k < 0 ? dflt : vals[k];
}

void ora_index_free(ora_index *ix) {
    if (!ix) return;
    for (int64_t i = 0; i < ix->n_el; i++) {
        oelem *e = &ix->el[i];
        for (uint32_t k = 0; k < e->nreads; k++) free(e->readnames[k]);
        free(e->readnames); free(e->name); free(e->fname); free(e->cname);
    }
    free(ix->el);
    for (int32_t i = 0; i < ix->chroms.n; i++) free(ix->bks[i].binHead);
    free(ix->bks);
    for (int32_t i = 0; i < ix->subs.n; i++) { free(ix->sub[i].bp_total); free(ix->sub[i].bp_total_unique); free(ix->sub[i].cpgScore); free(ix->sub[i].fname); free(ix->sub[i].cname); }
    free(ix->sub);
    for (int32_t i = 0; i < ix->fams.n; i++) free(ix->fam[i].cname);
    free(ix->fam); free(ix->cla);
    ktab_free(&ix->chromsize_names); free(ix->chromsize_val);
    ktab_free(&ix->repsize_names); free(ix->repsize_val);
    ktab_free(&ix->chroms); ktab_free(&ix->subs); ktab_free(&ix->fams); ktab_free(&ix->clas);
    ktab_free(&ix->nochr); ktab_free(&ix->dup);
    free(ix->sub_order); free(ix->fam_order); free(ix->cla_order);
    free(ix);
}

ora_index *ora_index_build(const char *chrom_sizes, const char *rep_sizes, const char *rmsk,
                           int filter_field, const char *filter_name, char err[256]) {
    ora_index *ix = calloc(1, sizeof *ix);
    if (err) err[0] = 0;
    ix->filter_field = filter_field;
    ktab_init(&ix->chroms, 0); ktab_init(&ix->subs, 0); ktab_init(&ix->fams, 0); ktab_init(&ix->clas, 0);
    ktab_init(&ix->nochr, 0); ktab_init(&ix->dup, 0);
    if (load_name_int(chrom_sizes, &ix->chromsize_names, &ix->chromsize_val, err)) { ora_index_free(ix); return NULL; }
    if (load_name_int(rep_sizes, &ix->repsize_names, &ix->repsize_val, err)) { ora_index_free(ix); return NULL; }
    FILE *f = fopen(rmsk, "r");
    if (!f) { if (err) snprintf(err, 256, "Couldn't open %s , %s", rmsk, strerror(errno)); ora_index_free(ix); return NULL; }
    char *line = NULL; size_t lc = 0; ssize_t n; long lineno = 0; int32_t row = -1; long repeat_num = 0;
    while ((n = getline(&line, &lc, f)) >= 0) {
        lineno++;
        if (line[0] == '#') continue;
        char *w[17]; int nw = chop_white(line, w, 17);
        if (nw == 0) continue;
        if (nw < 17) { if (err) snprintf(err, 256, "Expecting 17 words line %ld of %s got %d", lineno, rmsk, nw); goto fail; }
        row++;
        if (filter_field != 0 && strcmp(filter_name, w[filter_field]) != 0) continue;
        repeat_num++;
        oelem e; memset(&e, 0, sizeof e);
        char strand = w[9][0];
        e.cons_start = (unsigned int)strtol(strand == '+' ? w[13] : w[15], NULL, 0);
        e.cons_end = (unsigned int)strtol(w[14], NULL, 0);
        unsigned int us = (unsigned int)strtol(w[6], NULL, 0), ue = (unsigned int)strtol(w[7], NULL, 0);
        e.start = (int)us; e.end = (int)ue; e.length = ue - us; e.row = row; e.next = -1;
        /* chromosome -> binKeeper, created on first sight; rows on chromosomes absent from the size
         * file are dropped before any subfamily bookkeeping (generic.c:1613-1626) */
        int32_t c = ktab_find(&ix->chroms, w[5]);
        if (c < 0) {
            int size = name_int_default(&ix->chromsize_names, ix->chromsize_val, w[5], 0);
            if (size == 0) continue;
            if (size < 0) { if (err) snprintf(err, 256, "bad range %d,%d in binKeeperNew", 0, size); goto fail; }
            c = ktab_add(&ix->chroms, w[5]);
            if (c >= ix->bk_cap) { ix->bk_cap = ix->bk_cap ? ix->bk_cap * 2 : 32; ix->bks = realloc(ix->bks, sizeof(obk) * ix->bk_cap); }
            obk *bk = &ix->bks[c]; bk->minPos = 0; bk->maxPos = size;
            bk->binCount = bin_from_range(size - 1, size) + 1;
            bk->binHead = malloc(sizeof(int64_t) * bk->binCount);
            for (int i = 0; i < bk->binCount; i++) bk->binHead[i] = -1;
        }
        obk *bk = &ix->bks[c];
        if (e.start < bk->minPos || e.end > bk->maxPos || e.start > e.end) {
            if (err) snprintf(err, 256, "(%d %d) out of range (%d %d) in binKeeperAdd", e.start, e.end, bk->minPos, bk->maxPos);
            goto fail;
        }
        int bin = bin_from_range(e.start, e.end);
        if (bin < 0) { if (err) snprintf(err, 256, "start %d, end %d out of range in findBin (max is 2Gb)", e.start, e.end); goto fail; }
        e.chrom = c;
        e.name = strdup(w[10]); e.cname = strdup(w[11]); e.fname = strdup(w[12]);
        e.sub = e.fam = e.cla = -1;
        if (filter_field == 0) {
            int32_t s = ktab_find(&ix->subs, w[10]);
            if (s < 0) {
                s = ktab_add(&ix->subs, w[10]);
                if (s >= ix->sub_cap) { ix->sub_cap = ix->sub_cap ? ix->sub_cap * 2 : 256; ix->sub = realloc(ix->sub, sizeof(osub) * ix->sub_cap); }
                osub *S = &ix->sub[s]; memset(S, 0, sizeof *S);
                S->genome_count = 1; S->total_length = e.length;
                S->length = (uint32_t)name_int_default(&ix->repsize_names, ix->repsize_val, w[10], 0);
                S->bp_total = calloc(S->length ? S->length : 1, 4); S->bp_total_unique = calloc(S->length ? S->length : 1, 4);
                S->cpgScore = calloc(S->length ? S->length : 1, 8);
                S->fname = strdup(w[12]); S->cname = strdup(w[11]);
            } else { ix->sub[s].genome_count++; ix->sub[s].total_length += e.length; }
            int32_t fa = ktab_find(&ix->fams, w[12]);
            if (fa < 0) {
                fa = ktab_add(&ix->fams, w[12]);
                if (fa >= ix->fam_cap) { ix->fam_cap = ix->fam_cap ? ix->fam_cap * 2 : 64; ix->fam = realloc(ix->fam, sizeof(ofam) * ix->fam_cap); }
                memset(&ix->fam[fa], 0, sizeof(ofam)); ix->fam[fa].genome_count = 1; ix->fam[fa].total_length = e.length; ix->fam[fa].cname = strdup(w[11]);
            } else { ix->fam[fa].genome_count++; ix->fam[fa].total_length += e.length; }
            int32_t cl = ktab_find(&ix->clas, w[11]);
            if (cl < 0) {
                cl = ktab_add(&ix->clas, w[11]);
                if (cl >= ix->cla_cap) { ix->cla_cap = ix->cla_cap ? ix->cla_cap * 2 : 64; ix->cla = realloc(ix->cla, sizeof(ofam) * ix->cla_cap); }
                memset(&ix->cla[cl], 0, sizeof(ofam)); ix->cla[cl].genome_count = 1; ix->cla[cl].total_length = e.length;
            } else { ix->cla[cl].genome_count++; ix->cla[cl].total_length += e.length; }
            e.sub = s; e.fam = fa; e.cla = cl;
        }
        if (ix->n_el == ix->el_cap) { ix->el_cap = ix->el_cap ? ix->el_cap * 2 : 1024; ix->el = realloc(ix->el, sizeof(oelem) * ix->el_cap); }
        e.next = bk->binHead[bin]; bk->binHead[bin] = ix->n_el;       /* slAddHead */
        ix->el[ix->n_el++] = e;
    }
    free(line); fclose(f);
    if (filter_field != 0 && repeat_num <= 0) {
        if (err) snprintf(err, 256, "* No repeats found related to [%s], typo? or specify wrong repName/Class/Family filter?", filter_name);
        ora_index_free(ix); return NULL;
    }
    return ix;
fail:
    free(line); fclose(f); ora_index_free(ix); return NULL;
}

void ora_index_reset_counts(ora_index *ix) {
    for (int32_t i = 0; i < ix->subs.n; i++) {
        osub *S = &ix->sub[i]; S->read_count = S->read_count_unique = 0; S->cpgCount = 0; S->cpgTotalScore = 0;
        memset(S->bp_total, 0, 4 * (size_t)S->length); memset(S->bp_total_unique, 0, 4 * (size_t)S->length); memset(S->cpgScore, 0, 8 * (size_t)S->length);
    }
    for (int32_t i = 0; i < ix->fams.n; i++) { ix->fam[i].read_count = ix->fam[i].read_count_unique = 0; ix->fam[i].cpgCount = 0; ix->fam[i].cpgTotalScore = 0; }
    for (int32_t i = 0; i < ix->clas.n; i++) { ix->cla[i].read_count = ix->cla[i].read_count_unique = 0; ix->cla[i].cpgCount = 0; ix->cla[i].cpgTotalScore = 0; }
    for (int64_t i = 0; i < ix->n_el; i++) {
        oelem *e = &ix->el[i];
        for (uint32_t k = 0; k < e->nreads; k++) free(e->readnames[k]);
        free(e->readnames); e->readnames = NULL; e->readcap = 0; e->nreads = e->nreads_unique = 0; e->cpgCount = 0; e->cpgTotalScore = 0;
    }
    ktab_free(&ix->nochr); ktab_init(&ix->nochr, 0);
    ktab_free(&ix->dup); ktab_init(&ix->dup, 0);
    ix->dupkey[0] = 0;
    memset(ix->cnt, 0, sizeof ix->cnt);
}

/* binKeeperFind: returns hit element indices in RESULT-LIST order (head first). */
typedef struct { int64_t *v; int n, cap; } hitlist;
static void bk_find(const ora_index *ix, const obk *bk, int start, int end, hitlist *h) {
    h->n = 0;
    if (start < bk->minPos) start = bk->minPos;
    if (end > bk->maxPos) end = bk->maxPos;
    if (start >= end) return;
    int sb = start >> BIN_FIRST_SHIFT, eb = (end - 1) >> BIN_FIRST_SHIFT;
    /* collected in traversal order; the reference prepends each hit, so the list is the reverse */
    for (int i = 0; i < 6; i++) {
        int off = BIN_OFFSETS[i];
        for (int j = sb + off; j <= eb + off; j++)
            for (int64_t k = bk->binHead[j]; k >= 0; k = ix->el[k].next) {
                const oelem *e = &ix->el[k];
                int s = e->start > start ? e->start : start, t = e->end < end ? e->end : end;
                if (t - s > 0) {
                    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->v = realloc(h->v, sizeof(int64_t) * h->cap); }
                    h->v[h->n++] = k;
                }
            }
        sb >>= BIN_NEXT_SHIFT; eb >>= BIN_NEXT_SHIFT;
    }
    for (int a = 0, b = h->n - 1; a < b; a++, b--) { int64_t t = h->v[a]; h->v[a] = h->v[b]; h->v[b] = t; }
}

/* getCov (generic.c:296-301) */
static float get_cov(unsigned int aStart, unsigned int aEnd, unsigned int start, unsigned int end) {
    int s = (int)aStart > (int)start ? (int)aStart : (int)start;
    int e = (int)aEnd < (int)end ? (int)aEnd : (int)end;
    int r = e - s; if (r < 0) r = 0;
    float overlap = (float)r;
    float denominator = (float)(aEnd - aStart);
    return (denominator == 0) ? 0.0f : overlap / denominator;
}

/* selection loop (generic.c:950-970): returns position (1-based) in the list or 0, and tcoverage */
static int select_last_ascent(const ora_index *ix, const hitlist *h, unsigned int start, unsigned int end, float *tcov) {
    float coverage = 0.0f, tcoverage = 0.0f; int index = 0, tindex = 0;
    for (int i = 0; i < h->n; i++) {
        index++;
        const oelem *e = &ix->el[h->v[i]];
        float cov = get_cov(start, end, (unsigned int)e->start, (unsigned int)e->end);
        if (cov > coverage) { tindex = index; tcoverage = cov; }
        coverage = cov;
    }
    *tcov = tcoverage;
    return tindex;
}

static int same_word(const char *a, const char *b) { return strcasecmp(a, b) == 0; }

/* chopByChar (cuskent/common.c:2029-2053) */
static int chop_char(char *in, char ch, char **out, int max) {
    int i; char c;
    if (*in == 0) return 0;
    for (i = 0; i < max; i++) {
        out[i] = in;
        for (;;) { if ((c = *in++) == 0) return i + 1; else if (c == ch) { in[-1] = 0; break; } }
    }
    return i;
}

/* mapped2diffSubfam (generic.c:303-341) */
static int mapped_to_diff_subfam(const ora_index *ix, const char *subfam, int nm, char *ah, int qlen, hitlist *scratch) {
    char *row[100], *row2[4];
    int nf = chop_char(ah, ';', row, 100);
    for (int i = 0; i < nf; i++) {
        if (strlen(row[i]) == 0) continue;
        int n2 = chop_char(row[i], ',', row2, 4);
        if (n2 != 4) continue;                  /* the reference asserts; malformed XA is undefined there */
        int nm2 = (int)strtol(row2[3], 0, 0);
        if (nm2 > nm) continue;
        int start = abs((int)strtol(row2[1], 0, 0));
        int end = start + qlen;
        int32_t c = ktab_find(&ix->chroms, row2[0]);
        if (c < 0) continue;
        bk_find(ix, &ix->bks[c], start, end, scratch);
        for (int k = 0; k < scratch->n; k++)
            if (!same_word(ix->el[scratch->v[k]].name, subfam)) return 1;
    }
    return 0;
}

/* bam_aux_get (cussamtools/bam_aux.c:28-48) */
static const uint8_t *aux_get(const uint8_t *aux, const uint8_t *end, const char tag[2]) {
    const uint8_t *s = aux; int y = tag[0] << 8 | tag[1];
    while (s < end) {
        int x = (int)s[0] << 8 | s[1]; s += 2;
        if (x == y) return s;
        int type = toupper(*s); ++s;
        if (type == 'Z' || type == 'H') { while (s < end && *s) ++s; ++s; }
        else if (type == 'B') { int sub = *s; int sz = (sub == 'C' || sub == 'c' || sub == 'A') ? 1 : (sub == 'S' || sub == 's') ? 2 : (sub == 'I' || sub == 'i' || sub == 'f') ? 4 : 0; int32_t cnt; memcpy(&cnt, s + 1, 4); s += 5 + (size_t)sz * cnt; }
        else { /* the type was upper-cased first, so 'f' and 'd' values get size 0 -- faithful to bam_aux.c:28-34 + bam.h:754-760 */
            s += (type == 'C' || type == 'A') ? 1 : (type == 'S') ? 2 : (type == 'I') ? 4 : 0; }
    }
    return NULL;
}
/* bam_aux2i (bam_aux.c:159-170) */
static int32_t aux2i(const uint8_t *s) {
    if (!s) return 0;
    int type = *s++;
    if (type == 'c') return (int8_t)*s;
    if (type == 'C') return *s;
    if (type == 's') { int16_t v; memcpy(&v, s, 2); return v; }
    if (type == 'S') { uint16_t v; memcpy(&v, s, 2); return v; }
    if (type == 'i' || type == 'I') { int32_t v; memcpy(&v, s, 4); return v; }
    return 0;
}

static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

int ora_scan_bam_stream(ora_index *ix, const uint8_t *bam, uint64_t len, const ora_opts *o,
                        uint64_t cnt_out[13], ora_trace *trace, uint64_t trace_cap, uint64_t *n_records) {
    uint64_t p = 0, nrec = 0;
    if (len < 12 || memcmp(bam, "BAM\1", 4) != 0) return -1;
    int32_t l_text; memcpy(&l_text, bam + 4, 4); p = 8 + (uint64_t)l_text;
    int32_t n_ref; memcpy(&n_ref, bam + p, 4); p += 4;
    char **tname = calloc(n_ref > 0 ? n_ref : 1, sizeof(char *));
    for (int i = 0; i < n_ref; i++) { int32_t l; memcpy(&l, bam + p, 4); tname[i] = (char *)(bam + p + 4); p += 4 + (uint64_t)l + 4; }
    uint64_t *c = ix->cnt;
    hitlist H = {0}, H2 = {0};
    char chr[512], ah[2048];
    const unsigned mapQ = o->mapQ, ext = o->extension, iSize = o->iSize;
    while (p + 4 <= len) {
        int32_t block_len; memcpy(&block_len, bam + p, 4);
        if (p + 4 + 32 > len) break;                                /* truncated core -> loop ends */
        uint32_t x[8]; memcpy(x, bam + p + 4, 32);
        if (p + 4 + (uint64_t)(uint32_t)block_len > len) break;     /* truncated data */
        const uint8_t *data = bam + p + 36;
        p += 4 + (uint64_t)(uint32_t)block_len;
        int32_t tid = (int32_t)x[0], pos = (int32_t)x[1];
        uint32_t qual = x[2] >> 8 & 0xff, l_qname = x[2] & 0xff, flag = x[3] >> 16, n_cigar = x[3] & 0xffff;
        int32_t l_qseq = (int32_t)x[4], mpos = (int32_t)x[6], isize = (int32_t)x[7];
        const uint32_t *cigar_p = (const uint32_t *)(data + l_qname);
        const uint8_t *aux = data + l_qname + 4 * (size_t)n_cigar + (size_t)l_qseq + (size_t)((l_qseq + 1) / 2);
        const uint8_t *aux_end = data + (block_len - 32);
        const char *qname = (const char *)data;
        ora_trace T = {0, 0, tid, -1, 0};
        uint64_t rec_i = nrec++;
        int paired = flag & 1, read1 = flag & 64;
        int slot2 = paired && !read1 && !o->treat;
        unsigned int start = 0, end = 0, cend; char strand = '+';
        int uniq = qual >= mapQ;
        c[slot2 ? 1 : 0]++;
        if (flag & 4) goto next;
        c[slot2 ? 3 : 2]++;
        if (tid < 0 || tid >= n_ref) goto next;                     /* the reference would read out of bounds */
        snprintf(chr, sizeof chr, "%s", tname[tid]);
        if (o->addChr) {
            if (strncmp(tname[tid], "GL", 2) == 0) goto next;
            else if (same_word(tname[tid], "MT")) strcpy(chr, "chrM");
            else if (strncmp(tname[tid], "chr", 3) != 0) snprintf(chr, sizeof chr, "chr%s", tname[tid]);
        }
        if (ktab_find(&ix->nochr, chr) >= 0) goto next;
        cend = (unsigned int)(name_int_default(&ix->chromsize_names, ix->chromsize_val, chr, 2) - 1);
        if (cend == 1) { ktab_add(&ix->nochr, chr); goto next; }    /* warn once, discard */
        c[slot2 ? 5 : 4]++;
        {
            int se_like = 0;
            if (o->treat) se_like = 1;
            else if (paired) {
                if (!(flag & 8)) {
                    if (!read1) goto next;
                    if ((unsigned int)abs(isize) > iSize || isize == 0) goto next;
                    c[6]++; if (uniq) c[7]++;
                    if (isize > 0) { start = (unsigned int)pos; strand = '+'; int te = (int)(start + (unsigned int)isize); end = umin(cend, (unsigned int)te); }
                    else { start = (unsigned int)mpos; strand = '-'; int te = (int)(start - (unsigned int)isize); end = umin(cend, (unsigned int)te); }
                } else { if (o->discardWrongEnd) goto next; se_like = 1; }
            } else se_like = 1;
            if (se_like) {
                c[6]++; if (uniq) c[7]++;
                start = (unsigned int)pos;
                int tmpend;
                /* CIGAR words past the record's own end are not read (the reference would read whatever follows in its
                 * buffer: undefined); the device path clamps the same way */
                uint32_t n_cig_in = n_cigar;
                { const int64_t room = (int64_t)block_len - 32 - (int64_t)l_qname; if (room < 4 * (int64_t)n_cigar) n_cig_in = room > 0 ? (uint32_t)(room / 4) : 0; }
                if (n_cigar) { uint32_t e2 = (uint32_t)pos; for (uint32_t k = 0; k < n_cig_in; k++) { uint32_t cg; memcpy(&cg, cigar_p + k, 4); int op = cg & 0xf; if (op == 0 || op == 2 || op == 3) e2 += cg >> 4; } tmpend = (int)e2; }
                else tmpend = pos + l_qseq;
                end = umin(cend, (unsigned int)tmpend);
                strand = (flag & 16) ? '-' : '+';
                if (ext) { if (strand == '+') end = umin(start + ext, cend); else start = (end < ext) ? 0 : end - ext; }
            }
        }
        if (o->rmDup) {
            if (uniq) snprintf(ix->dupkey, sizeof ix->dupkey, "%s:%u:%u:%c", chr, start, end, strand);
            if (ktab_find(&ix->dup, ix->dupkey) >= 0) goto next;
            ktab_add(&ix->dup, ix->dupkey);
        }
        if (uniq) c[11]++;
        T.start = start; T.end = end; T.flags = ORA_T_FRAGMENT | (uniq ? ORA_T_UNIQ : 0) | (strand == '-' ? ORA_T_MINUS : 0);
        {
            unsigned int qlen = end - start;
            int32_t ci = ktab_find(&ix->chroms, chr);
            const uint8_t *xa = aux_get(aux, aux_end, "XA");
            if (xa) T.flags |= ORA_T_HAS_XA;
            if (ci < 0) goto next;
            bk_find(ix, &ix->bks[ci], (int)start, (int)end, &H);
            if (H.n == 0) goto next;
            float tcov; int tindex = select_last_ascent(ix, &H, start, end, &tcov);
            if (tcov < o->minCoverage) goto next;
            if (tindex == 0) goto next;                              /* -c 0 with zero coverage: the reference dereferences NULL */
            oelem *ss = &ix->el[H.v[tindex - 1]];
            T.sel_row = ss->row;
            if (o->diffSubfam && xa) {
                if (*xa == 'Z' || *xa == 'H') {
                    snprintf(ah, sizeof ah, "%s", (const char *)(xa + 1));
                    int nm = aux2i(aux_get(aux, aux_end, "NM"));
                    if (mapped_to_diff_subfam(ix, ss->name, nm, ah, (int)qlen, &H2)) { c[12]++; T.flags |= ORA_T_DIFFSUB; goto next; }
                }
            }
            if (o->filter == 0) {
                if (ss->sub >= 0) {
                    osub *rs = &ix->sub[ss->sub];
                    rs->read_count++; if (uniq) rs->read_count_unique++;
                    if (rs->length != 0) {
                        unsigned int rstart = start - (unsigned int)ss->start;
                        unsigned int rend = rstart + qlen;
                        rend = (rend < (unsigned int)ss->end) ? rend : (unsigned int)ss->end;
                        for (int i = (int)rstart; (unsigned int)i < rend; i++) {
                            int j = (int)((unsigned int)i + ss->cons_start);
                            if ((unsigned int)j >= ss->cons_end) break;
                            if ((unsigned int)j >= rs->length) break;
                            rs->bp_total[j]++; if (uniq) rs->bp_total_unique[j]++;
                        }
                    }
                }
                if (ss->fam >= 0) { ix->fam[ss->fam].read_count++; if (uniq) ix->fam[ss->fam].read_count_unique++; }
                if (ss->cla >= 0) { ix->cla[ss->cla].read_count++; if (uniq) ix->cla[ss->cla].read_count_unique++; }
            } else {
                if (ss->nreads == ss->readcap) { ss->readcap = ss->readcap ? ss->readcap * 2 : 4; ss->readnames = realloc(ss->readnames, sizeof(char *) * ss->readcap); }
                ss->readnames[ss->nreads++] = strdup(qname);
                if (uniq) ss->nreads_unique++;
            }
            c[9]++; if (uniq) c[10]++;
            T.flags |= ORA_T_COUNTED;
        }
    next:
        if (trace && rec_i < trace_cap) trace[rec_i] = T;
    }
    free(tname); free(H.v); free(H2.v);
    if (cnt_out) memcpy(cnt_out, c, sizeof(uint64_t) * 13);
    if (n_records) *n_records = nrec;
    return 0;
}

/* BGZF inflate the way the reference reads it (bgzf.c:401-411 header check, 471-521 block read,
 * 524-565: an empty block, or EOF, ends the stream; no CRC check). */
uint8_t *ora_inflate_bam(const char *path, uint64_t *len) {
    FILE *f = fopen(path, "rb"); if (!f) return NULL;
    uint64_t cap = 1 << 24, n = 0; uint8_t *out = malloc(cap);
    uint8_t *cb = malloc(65536 + 64);
    for (;;) {
        size_t got = fread(cb, 1, 18, f);
        if (got == 0) break;
        if (got != 18) break;
        if (!(cb[0] == 31 && cb[1] == 139 && cb[2] == 8 && (cb[3] & 4) && cb[10] == 6 && cb[11] == 0 && cb[12] == 'B' && cb[13] == 'C' && cb[14] == 2 && cb[15] == 0)) break;
        int block_length = (cb[16] | cb[17] << 8) + 1;
        int remaining = block_length - 18;
        if ((int)fread(cb + 18, 1, remaining, f) != remaining) break;
        if (n + 65536 > cap) { cap *= 2; out = realloc(out, cap); }
        z_stream zs; memset(&zs, 0, sizeof zs);
        zs.next_in = cb + 18; zs.avail_in = block_length - 16; zs.next_out = out + n; zs.avail_out = 65536;
        if (inflateInit2(&zs, -15) != Z_OK) break;
        int st = inflate(&zs, Z_FINISH); inflateEnd(&zs);
        if (st != Z_STREAM_END) break;
        if (zs.total_out == 0) break;                                /* empty block ends the stream */
        n += zs.total_out;
    }
    free(cb); fclose(f);
    *len = n;
    return out;
}
void ora_free(void *p) { free(p); }

int ora_scan_bam_file(ora_index *ix, const char *path, const ora_opts *o, uint64_t cnt[13]) {
    uint64_t len; uint8_t *b = ora_inflate_bam(path, &len);
    if (!b) return -1;
    int rc = ora_scan_bam_stream(ix, b, len, o, cnt, NULL, 0, NULL);
    free(b);
    return rc;
}

int ora_scan_cpg(ora_index *ix, const char *bedgraph, int filter, uint32_t *cpg_lines, uint32_t *cpg_in_repeat, char err[256]) {
    FILE *f = fopen(bedgraph, "r");
    if (!f) { if (err) snprintf(err, 256, "Couldn't open %s , %s", bedgraph, strerror(errno)); return -1; }
    char *line = NULL; size_t lc = 0; ssize_t n; unsigned int lines = 0, inrep = 0;
    hitlist H = {0};
    while ((n = getline(&line, &lc, f)) >= 0) {
        char *s = line; while (isspace((unsigned char)*s)) s++;
        if (*s == 0 || *s == '#') continue;
        char *row[20]; int nf = chop_white(line, row, 20);
        if (nf < 4) { if (err) snprintf(err, 256, "file %s doesn't appear to be in bedGraph format. At least 4 fields required, got %d", bedgraph, nf); free(line); fclose(f); free(H.v); return -1; }
        lines++;
        unsigned int start = (unsigned int)strtol(row[1], NULL, 0), end = (unsigned int)strtol(row[2], NULL, 0);
        double score = strtod(row[3], NULL);
        int32_t ci = ktab_find(&ix->chroms, row[0]);
        if (ci < 0) continue;
        bk_find(ix, &ix->bks[ci], (int)start, (int)end, &H);
        if (H.n == 0) continue;
        oelem *ss = &ix->el[H.v[0]];
        if (filter) { ss->cpgCount++; ss->cpgTotalScore += score; }
        else {
            if (ss->sub >= 0) {
                osub *rs = &ix->sub[ss->sub];
                rs->cpgCount++; rs->cpgTotalScore += score;
                if (rs->length != 0) {
                    unsigned int rstart = start - (unsigned int)ss->start;
                    unsigned int rend = rstart + 2;
                    rend = (rend < (unsigned int)ss->end) ? rend : (unsigned int)ss->end;
                    for (int i = (int)rstart; (unsigned int)i < rend; i++) {
                        int j = (int)((unsigned int)i + ss->cons_start);
                        if ((unsigned int)j >= ss->cons_end) break;
                        if ((unsigned int)j >= rs->length) break;
                        rs->cpgScore[j] += score;
                    }
                }
            }
            if (ss->fam >= 0) { ix->fam[ss->fam].cpgCount++; ix->fam[ss->fam].cpgTotalScore += score; }
            if (ss->cla >= 0) { ix->cla[ss->cla].cpgCount++; ix->cla[ss->cla].cpgTotalScore += score; }
        }
        inrep++;
    }
    free(line); fclose(f); free(H.v);
    if (cpg_lines) *cpg_lines = lines;
    if (cpg_in_repeat) *cpg_in_repeat = inrep;
    return 0;
}

/* cal_rpkm / cal_rpm (generic.c:35-41): same operation order */
static double cal_rpkm(unsigned long long reads_count, unsigned long long total_length, unsigned long long mapped) { return reads_count / (mapped * 1e-9 * total_length); }
static double cal_rpm(unsigned long long reads_count, unsigned long long mapped) { return reads_count / (mapped * 1e-6); }

static void orders(ora_index *ix) {
    if (!ix->sub_order) ix->sub_order = ktab_order(&ix->subs);
    if (!ix->fam_order) ix->fam_order = ktab_order(&ix->fams);
    if (!ix->cla_order) ix->cla_order = ktab_order(&ix->clas);
}

int ora_write_stat(ora_index *ix, const char *of1, const char *of2, const char *of3, const char *of4, const char *of5,
                   uint64_t reads_num, uint64_t reads_num_unique) {
    orders(ix);
    FILE *f1 = fopen(of1, "w"), *f2 = fopen(of2, "w"), *f5 = fopen(of5, "w");
    if (!f1 || !f2 || !f5) return -1;
    fprintf(f1, "#subfamily\tfamily\tclass\tconsensus_length\treads_count\tunique_reads_count\ttotal_length\tgenome_count\tall_reads_RPKM\tall_reads_RPM\tunique_reads_RPKM\tunique_reads_RPM\n");
    for (int32_t k = 0; k < ix->subs.n; k++) {
        int32_t i = ix->sub_order[k]; osub *S = &ix->sub[i];
        fprintf(f1, "%s\t%s\t%s\t%u\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->subs.names[i], S->fname, S->cname, S->length,
                (unsigned long long)S->read_count, (unsigned long long)S->read_count_unique, (unsigned long long)S->total_length, (unsigned long long)S->genome_count,
                cal_rpkm(S->read_count, S->total_length, reads_num), cal_rpm(S->read_count, reads_num),
                cal_rpkm(S->read_count_unique, S->total_length, reads_num_unique), cal_rpm(S->read_count_unique, reads_num_unique));
        if (S->length != 0) {
            fprintf(f2, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[i]);
            fprintf(f5, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[i]);
            for (uint32_t m = 0; m < S->length; m++) { fprintf(f2, "%u\n", S->bp_total[m]); fprintf(f5, "%u\n", S->bp_total_unique[m]); }
        }
    }
    fclose(f2); fclose(f1); fclose(f5);
    FILE *f3 = fopen(of3, "w"); if (!f3) return -1;
    fprintf(f3, "#family\tclass\treads_count\tunique_reads_count\ttotal_length\tgenome_count\tall_reads_RPKM\tall_reads_RPM\tunique_reads_RPKM\tunique_reads_RPM\n");
    for (int32_t k = 0; k < ix->fams.n; k++) {
        int32_t i = ix->fam_order[k]; ofam *S = &ix->fam[i];
        fprintf(f3, "%s\t%s\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->fams.names[i], S->cname,
                (unsigned long long)S->read_count, (unsigned long long)S->read_count_unique, (unsigned long long)S->total_length, (unsigned long long)S->genome_count,
                cal_rpkm(S->read_count, S->total_length, reads_num), cal_rpm(S->read_count, reads_num),
                cal_rpkm(S->read_count_unique, S->total_length, reads_num_unique), cal_rpm(S->read_count_unique, reads_num_unique));
    }
    fclose(f3);
    FILE *f4 = fopen(of4, "w"); if (!f4) return -1;
    fprintf(f4, "#class\treads_count\tunique_reads_count\ttotal_length\tgenome_count\tall_reads_RPKM\tall_reads_RPM\tunique_reads_RPKM\tunique_reads_RPM\n");
    for (int32_t k = 0; k < ix->clas.n; k++) {
        int32_t i = ix->cla_order[k]; ofam *S = &ix->cla[i];
        fprintf(f4, "%s\t%llu\t%llu\t%llu\t%llu\t%.3f\t%.3f\t%.3f\t%.3f\n", ix->clas.names[i],
                (unsigned long long)S->read_count, (unsigned long long)S->read_count_unique, (unsigned long long)S->total_length, (unsigned long long)S->genome_count,
                cal_rpkm(S->read_count, S->total_length, reads_num), cal_rpm(S->read_count, reads_num),
                cal_rpkm(S->read_count_unique, S->total_length, reads_num_unique), cal_rpm(S->read_count_unique, reads_num_unique));
    }
    fclose(f4);
    return 0;
}

int ora_write_report(const char *path, const uint64_t cnt[13], uint32_t mapQ, const char *subfam) {
    FILE *f = fopen(path, "w"); if (!f) return -1;
    fprintf(f, "total reads (pair): %llu\n", (unsigned long long)cnt[0]);
    fprintf(f, "mappable reads (pair): %llu\n", (unsigned long long)cnt[6]);
    fprintf(f, "uniquely mapped reads (pair) (mapQ >= %u): %llu\n", mapQ, (unsigned long long)cnt[7]);
    fprintf(f, "non-redundant uniquely mapped reads (pair): %llu\n", (unsigned long long)cnt[11]);
    fprintf(f, "mapped reads (pair) overlap with repeats but discarded due to mapped to different subfamilies: %llu\n", (unsigned long long)cnt[12]);
    fprintf(f, "mapped reads (pair) overlap with [%s] repeats: %llu\n", subfam, (unsigned long long)cnt[9]);
    fprintf(f, "uniquely mapped reads (pair) overlap with [%s] repeats: %llu\n", subfam, (unsigned long long)cnt[10]);
    fclose(f);
    return 0;
}

/* writeFilterOut (generic.c:1709-1746): chromosome hash order, then binKeeperFirst/Next (bins
 * ascending, newest first inside a bin). */
int ora_write_filter(ora_index *ix, const char *path, int readlist, int threshold, uint64_t reads_num) {
    FILE *f = fopen(path, "w"); if (!f) return -1;
    if (readlist) fprintf(f, "#chr\tstart\tend\tlength\trepName\trepClass\trepFamily\treadsCount\tRPKM\tRPM\treadsList\n");
    else fprintf(f, "#chr\tstart\tend\tlength\trepName\trepClass\trepFamily\treadsCount\tRPKM\tRPM\n");
    int32_t *co = ktab_order(&ix->chroms);
    for (int32_t k = 0; k < ix->chroms.n; k++) {
        int32_t c = co[k]; obk *bk = &ix->bks[c];
        for (int b = 0; b < bk->binCount; b++)
            for (int64_t e = bk->binHead[b]; e >= 0; e = ix->el[e].next) {
                oelem *os = &ix->el[e];
                int count = (int)os->nreads;
                if (count < threshold) continue;
                fprintf(f, "%s\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%.3f\t%.3f", ix->chroms.names[c], os->start, os->end, (int)os->length, os->name, os->cname, os->fname, count,
                        cal_rpkm((unsigned long long)count, (unsigned long long)os->length, reads_num), cal_rpm((unsigned long long)count, reads_num));
                if (readlist) { fputc('\t', f); for (uint32_t r = 0; r < os->nreads; r++) { if (r) fputc(',', f); fputs(os->readnames[r], f); } }
                fputc('\n', f);
            }
    }
    free(co); fclose(f);
    return 0;
}

int ora_write_cpg_stat(ora_index *ix, const char *of1, const char *of2, const char *of3, const char *of4) {
    orders(ix);
    FILE *f1 = fopen(of1, "w"), *f2 = fopen(of2, "w"); if (!f1 || !f2) return -1;
    fprintf(f1, "#subfamily\tfamily\tclass\tconsensus_length\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->subs.n; k++) {
        int32_t i = ix->sub_order[k]; osub *S = &ix->sub[i];
        fprintf(f1, "%s\t%s\t%s\t%u\t%u\t%.4f\t%llu\t%llu\n", ix->subs.names[i], S->fname, S->cname, S->length, S->cpgCount, S->cpgTotalScore, (unsigned long long)S->total_length, (unsigned long long)S->genome_count);
        if (S->length != 0) { fprintf(f2, "fixedStep chrom=%s start=1 step=1 span=1\n", ix->subs.names[i]); for (uint32_t m = 0; m < S->length; m++) fprintf(f2, "%.4f\n", S->cpgScore[m]); }
    }
    fclose(f2); fclose(f1);
    FILE *f3 = fopen(of3, "w"); if (!f3) return -1;
    fprintf(f3, "#family\tclass\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->fams.n; k++) { int32_t i = ix->fam_order[k]; ofam *S = &ix->fam[i];
        fprintf(f3, "%s\t%s\t%u\t%.4f\t%llu\t%llu\n", ix->fams.names[i], S->cname, S->cpgCount, S->cpgTotalScore, (unsigned long long)S->total_length, (unsigned long long)S->genome_count); }
    fclose(f3);
    FILE *f4 = fopen(of4, "w"); if (!f4) return -1;
    fprintf(f4, "#class\tcovered_CpG_sites\tCpG_total_score\ttotal_length\tgenome_count\n");
    for (int32_t k = 0; k < ix->clas.n; k++) { int32_t i = ix->cla_order[k]; ofam *S = &ix->cla[i];
        fprintf(f4, "%s\t%u\t%.4f\t%llu\t%llu\n", ix->clas.names[i], S->cpgCount, S->cpgTotalScore, (unsigned long long)S->total_length, (unsigned long long)S->genome_count); }
    fclose(f4);
    return 0;
}

int ora_write_cpg_filter(ora_index *ix, const char *path, double thr) {
    FILE *f = fopen(path, "w"); if (!f) return -1;
    fprintf(f, "#chr\tstart\tend\tlength\trepName\trepClass\trepFamily\tcovered_CpG_site\ttotal_CpG_score\n");
    int32_t *co = ktab_order(&ix->chroms);
    for (int32_t k = 0; k < ix->chroms.n; k++) {
        int32_t c = co[k]; obk *bk = &ix->bks[c];
        for (int b = 0; b < bk->binCount; b++)
            for (int64_t e = bk->binHead[b]; e >= 0; e = ix->el[e].next) {
                oelem *os = &ix->el[e];
                if (os->cpgTotalScore > thr)
                    fprintf(f, "%s\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%.3f\n", ix->chroms.names[c], os->start, os->end, (int)os->length, os->name, os->cname, os->fname, (int)os->cpgCount, os->cpgTotalScore);
            }
    }
    free(co); fclose(f);
    return 0;
}

int32_t ora_find_select(ora_index *ix, const char *chrom, uint32_t start, uint32_t end, float min_cov,
                        int32_t *n_hits, int32_t *hits, int32_t cap) {
    if (n_hits) *n_hits = 0;
    int32_t ci = ktab_find(&ix->chroms, chrom); if (ci < 0) return -1;
    hitlist H = {0};
    bk_find(ix, &ix->bks[ci], (int)start, (int)end, &H);
    if (n_hits) *n_hits = H.n;
    for (int i = 0; i < H.n && i < cap; i++) hits[i] = ix->el[H.v[i]].row;
    int32_t sel = -1;
    if (H.n) { float tc; int t = select_last_ascent(ix, &H, start, end, &tc); if (!(tc < min_cov) && t > 0) sel = ix->el[H.v[t - 1]].row; }
    free(H.v);
    return sel;
}

int32_t ora_n_subfam(ora_index *ix) { return ix->subs.n; }
int32_t ora_n_fam(ora_index *ix) { return ix->fams.n; }
int32_t ora_n_class(ora_index *ix) { return ix->clas.n; }
int64_t ora_n_elem(ora_index *ix) { return ix->n_el; }
const char *ora_name(ora_index *ix, int which, int32_t i) {
    orders(ix);
    if (which == 0) return ix->subs.names[ix->sub_order[i]];
    if (which == 1) return ix->fams.names[ix->fam_order[i]];
    return ix->clas.names[ix->cla_order[i]];
}
void ora_counts(ora_index *ix, int which, int32_t i, uint64_t out[4]) {
    orders(ix);
    if (which == 0) { osub *S = &ix->sub[ix->sub_order[i]]; out[0] = S->read_count; out[1] = S->read_count_unique; out[2] = S->total_length; out[3] = S->genome_count; }
    else { ofam *S = which == 1 ? &ix->fam[ix->fam_order[i]] : &ix->cla[ix->cla_order[i]]; out[0] = S->read_count; out[1] = S->read_count_unique; out[2] = S->total_length; out[3] = S->genome_count; }
}
uint32_t ora_subfam_length(ora_index *ix, int32_t i) { orders(ix); return ix->sub[ix->sub_order[i]].length; }
const uint32_t *ora_subfam_bp(ora_index *ix, int32_t i, int unique) { orders(ix); osub *S = &ix->sub[ix->sub_order[i]]; return unique ? S->bp_total_unique : S->bp_total; }
