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

int itx_host_index_load(struct itx_index *ix, const char *chrom_sizes, const char *rep_sizes, const char *rmsk,
                        int filter_field, const char *filter_name, char err[ITX_ERRLEN]) {
    int rc;
    ix->filter_field = filter_field; ix->stat_mode = (filter_field == 0);
    itx_strtab_init(&ix->chroms); itx_strtab_init(&ix->subs); itx_strtab_init(&ix->fams); itx_strtab_init(&ix->clas);
    itx_strtab_init(&ix->warned);
    if ((rc = load_name_int(chrom_sizes, &ix->chromsize, &ix->chromsize_val, err))) return rc;
    if ((rc = load_name_int(rep_sizes, &ix->repsize, &ix->repsize_val, err))) return rc;
    if (filter_field != 0 && (filter_field < 0 || filter_field > 16 || !filter_name)) { snprintf(err, ITX_ERRLEN, "bad filter field %d", filter_field); return ITX_EARG; }

    FILE *f = is_directory(rmsk) ? NULL : fopen(rmsk, "r");
    if (!f) { snprintf(err, ITX_ERRLEN, "Couldn't open %s , %s", rmsk, strerror(errno)); return ITX_EIO; }
    static const size_t IOBUF = 1 << 22; char *iobuf = (char *)malloc(IOBUF); setvbuf(f, iobuf, _IOFBF, IOBUF);
    raw_el *raw = NULL; long long nraw = 0, rawcap = 0;
    int32_t subcap = 0, famcap = 0, clacap = 0, chromcap = 0;
    long long row = -1, kept = 0; long ln = 0;
    char *line = NULL; size_t lc = 0;
    rc = ITX_OK;
    while (getline(&line, &lc, f) >= 0) {
        ln++;
        if (line[0] == '#') continue;
        char *w[17]; int nw = split_white(line, w, 17);
        if (nw == 0) continue;
        if (nw < 17) { snprintf(err, ITX_ERRLEN, "Expecting 17 words line %ld of %s got %d", ln, rmsk, nw); rc = ITX_EFORMAT; break; }
        row++;
        if (filter_field != 0 && strcmp(filter_name, w[filter_field]) != 0) continue;
        kept++;
        raw_el e;
        e.cs = (uint32_t)strtol(w[9][0] == '+' ? w[13] : w[15], NULL, 0);
        e.ce = (uint32_t)strtol(w[14], NULL, 0);
        e.start = (int32_t)(uint32_t)strtol(w[6], NULL, 0);
        e.end = (int32_t)(uint32_t)strtol(w[7], NULL, 0);
        e.row = (uint32_t)row;
        int32_t c = itx_strtab_find(&ix->chroms, w[5]);
        if (c < 0) {
            int size = name_int_or(&ix->chromsize, ix->chromsize_val, w[5], 0);
            if (size == 0) continue;                       /* chromosome not in the size file: row dropped */
            if (size < 0) { snprintf(err, ITX_ERRLEN, "bad range %d,%d in binKeeperNew", 0, size); rc = ITX_EFORMAT; break; }
            c = itx_strtab_add(&ix->chroms, w[5]);
            if (c >= chromcap) { chromcap = chromcap ? chromcap * 2 : 64; ix->chrom_size = (int32_t *)realloc(ix->chrom_size, sizeof(int32_t) * (size_t)chromcap); }
            ix->chrom_size[c] = size;
        }
        int32_t bin;
        if (e.start < 0 || e.end > ix->chrom_size[c] || e.start > e.end) {
            snprintf(err, ITX_ERRLEN, "(%d %d) out of range (%d %d) in binKeeperAdd", e.start, e.end, 0, ix->chrom_size[c]); rc = ITX_EFORMAT; break;
        }
        if (bin_level(e.start, e.end, &bin) < 0) { snprintf(err, ITX_ERRLEN, "start %d, end %d out of range in findBin (max is 2Gb)", e.start, e.end); rc = ITX_EFORMAT; break; }
        e.chrom = c;
        int32_t ns0 = ix->subs.n, nf0 = ix->fams.n, nc0 = ix->clas.n;
        e.sub = itx_strtab_intern(&ix->subs, w[10]);
        e.cla = itx_strtab_intern(&ix->clas, w[11]);
        e.fam = itx_strtab_intern(&ix->fams, w[12]);
        ix->sub = grow_groups(ix->sub, &subcap, ix->subs.n);
        ix->fam = grow_groups(ix->fam, &famcap, ix->fams.n);
        ix->cla = grow_groups(ix->cla, &clacap, ix->clas.n);
        if (ix->subs.n != ns0) { ix->sub[e.sub].first_fam = e.fam; ix->sub[e.sub].first_cla = e.cla; }
        if (ix->fams.n != nf0) { ix->fam[e.fam].first_cla = e.cla; }
        (void)nc0;
        if (ix->stat_mode) {
            uint64_t len = (uint32_t)(e.end - e.start);
            ix->sub[e.sub].genome_count++; ix->sub[e.sub].total_length += len;
            ix->fam[e.fam].genome_count++; ix->fam[e.fam].total_length += len;
            ix->cla[e.cla].genome_count++; ix->cla[e.cla].total_length += len;
        }
        if (nraw == rawcap) { rawcap = rawcap ? rawcap * 2 : (1 << 16); raw = (raw_el *)realloc(raw, sizeof(raw_el) * (size_t)rawcap); }
        raw[nraw++] = e;
    }
    free(line); fclose(f); free(iobuf);
    if (rc) { free(raw); return rc; }
    if (filter_field != 0 && kept <= 0) {
        snprintf(err, ITX_ERRLEN, "* No repeats found related to [%s], typo? or specify wrong repName/Class/Family filter?", filter_name);
        free(raw); return ITX_EFORMAT;
    }
    ix->n_rows = row + 1; ix->n_elem = nraw;

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
    if (ix->n_bucket >= 0xffffffffLL) { snprintf(err, ITX_ERRLEN, "genome too large for the position buckets"); return ITX_ENOTSUP; }
    ix->cinfo = (itx_chrominfo *)calloc((size_t)(nchrom ? nchrom : 1), sizeof(itx_chrominfo));
    for (int32_t c = 0; c < nchrom; c++) {
        ix->cinfo[c].off = (uint32_t)ix->chrom_off[c]; ix->cinfo[c].bucket = (uint32_t)ix->chrom_bucket[c];
        ix->cinfo[c].size = ix->chrom_size[c]; ix->cinfo[c].n = (uint32_t)(ix->chrom_off[c + 1] - ix->chrom_off[c]);
    }
    itx_strtab_free(&fold);
    return ITX_OK;
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
    free(ix->iv); free(ix->bucket); free(ix->chrom_bucket); free(ix->cinfo); free(ix->sinfo); free(ix->meta); free(ix->meta2); free(ix->el_chrom); free(ix->row2el);
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
            for (uint32_t m = 0; m < L; m++) { fprintf(fw, "%u\n", a[m]); fprintf(fu, "%u\n", u[m]); }
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

/* ------------------------------------------------------------------ accessors */
int32_t itx_n_subfam(const itx_index *ix) { return ix->stat_mode ? ix->subs.n : 0; }
int32_t itx_n_fam(const itx_index *ix) { return ix->stat_mode ? ix->fams.n : 0; }
int32_t itx_n_class(const itx_index *ix) { return ix->stat_mode ? ix->clas.n : 0; }
int64_t itx_n_elem(const itx_index *ix) { return ix->n_elem; }
int64_t itx_n_rows(const itx_index *ix) { return ix->n_rows; }
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
