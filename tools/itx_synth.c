/* itx_synth -- deterministic synthetic inputs for the iteres hot path (TEST / BENCH INFRASTRUCTURE).
 *
 * The reference ships no data, no fixtures and no generator; the shapes below are
 * the ones SURVEY.md 8(d) and BASELINE.json name.  Everything is a pure function of (seed, sizes),
 * independent of the thread count, so the same bytes are produced here, on the GPU box, and for
 * every rank's shard.
 *
 *   chrom sizes   shape 0: chr1 only;  shape 1: hg19 (chr1-22,X,Y,M).  The BAM header additionally
 *                 names one contig that is ABSENT from the size file (reads on it are discarded
 *                 with a warning by the reference, generic.c:796-801).
 *   repeat sizes  n_subfam names SUB0000.., consensus length U[100,6500]; 5 % omitted (length-0 path).
 *   rmsk.txt      17 tab-separated UCSC columns, rows sorted by start inside each chromosome,
 *                 ~2 % nested/overlapping rows, natural 128-kb bin straddlers, Zipf-like subfamily
 *                 usage, 0.5 % rows whose family/class differ from the subfamily's first row.
 *   BAM           mode 0: SE-50 (cfg1/2)   mode 1: SE-75 with NM/XA multi-reads (cfg3)
 *                 mode 2: PE-100 (cfg5).  Uncompressed record stream generated in parallel in
 *                 fixed chunks; BGZF writer (zlib raw deflate, <=0xff00-byte payloads, EOF block).
 *   bedGraph      cfg4: chrom start start+2 score (two decimals), sorted.
 *
 * Built as tools/libitx_synth.so (ctypes from tests/ and bench.py) and tools/itx_synth (CLI).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <zlib.h>

/* ------------------------------------------------------------------ RNG */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline uint64_t rng_next(rng_t *r) { r->s += 0x9e3779b97f4a7c15ULL; return mix64(r->s); }
static inline double rng_unif(rng_t *r) { return (rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return n ? (uint64_t)(rng_unif(r) * (double)n) : 0; }
static inline double rng_norm(rng_t *r) {
    double u1 = rng_unif(r), u2 = rng_unif(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static inline rng_t rng_seed(uint64_t seed, uint64_t stream) {
    rng_t r; r.s = mix64(seed * 0x100000001b3ULL + stream) ^ (stream << 17); return r;
}

/* ------------------------------------------------------------------ genome shapes */
static const char *HG19_NAMES[25] = {"chr1","chr2","chr3","chr4","chr5","chr6","chr7","chr8","chr9","chr10",
    "chr11","chr12","chr13","chr14","chr15","chr16","chr17","chr18","chr19","chr20","chr21","chr22","chrX","chrY","chrM"};
static const uint32_t HG19_LENS[25] = {249250621,243199373,198022430,191154276,180915260,171115067,159138663,
    146364022,141213431,135534747,135006516,133851895,115169878,107349540,102531392,90354753,81195210,78077248,
    59128983,63025520,48129895,51304566,155270560,59373566,16571};
#define ABSENT_NAME "chrUn_gl000220"
#define ABSENT_LEN 161802u

typedef struct {
    uint32_t chrom, start, end;       /* genomic, 0-based half-open */
    uint32_t subfam, fam, cla;
    uint32_t cs, ce;                  /* consensus window */
    uint32_t conslen;                 /* notional consensus length of the subfamily */
    char strand;
} srow_t;

typedef struct {
    int shape, n_chrom;
    const char **names; const uint32_t *lens;
    uint64_t cum[26];                 /* cumulative chromosome starts in linear coordinates */
    uint64_t glen;
    int n_subfam, n_fam, n_cla;
    uint32_t *conslen;                /* per subfamily */
    uint8_t *in_sizefile;             /* per subfamily: listed in the repeat-size file? */
    uint64_t n_rmsk;
    srow_t *rows;                     /* sorted by (chrom, start) */
    uint64_t chrom_row0[26];          /* first row of each chromosome */
    /* per-subfamily row lists (for XA alternates) */
    uint64_t *sf_off; uint64_t *sf_rows;
    uint64_t seed;
} synth_t;

static const char *CHR1_NAMES[1] = {"chr1"};
static const uint32_t CHR1_LENS[1] = {249250621};
/* shape 2: two small chromosomes, so that a few tens of thousands of reads pile up on the same coordinates (-R tests) */
static const char *SMALL_NAMES[2] = {"chr1", "chrM"};
static const uint32_t SMALL_LENS[2] = {400000, 16571};

static int cmp_row(const void *a, const void *b) {
    const srow_t *x = a, *y = b;
    if (x->chrom != y->chrom) return x->chrom < y->chrom ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    return 0;
}

void *synth_new(int shape, uint64_t n_rmsk, int n_subfam, int n_fam, int n_cla, uint64_t seed) {
    synth_t *S = calloc(1, sizeof(*S));
    S->shape = shape; S->seed = seed;
    if (shape == 0) { S->n_chrom = 1; S->names = CHR1_NAMES; S->lens = CHR1_LENS; }
    else if (shape == 2) { S->n_chrom = 2; S->names = SMALL_NAMES; S->lens = SMALL_LENS; }
    else { S->n_chrom = 25; S->names = HG19_NAMES; S->lens = HG19_LENS; }
    S->cum[0] = 0;
    for (int i = 0; i < S->n_chrom; i++) S->cum[i + 1] = S->cum[i] + S->lens[i];
    S->glen = S->cum[S->n_chrom];
    S->n_subfam = n_subfam; S->n_fam = n_fam; S->n_cla = n_cla;
    S->conslen = malloc(sizeof(uint32_t) * n_subfam);
    S->in_sizefile = malloc(n_subfam);
    rng_t r = rng_seed(seed, 0x5151);
    for (int i = 0; i < n_subfam; i++) {
        S->conslen[i] = 100 + (uint32_t)rng_below(&r, 6401);
        S->in_sizefile[i] = rng_unif(&r) >= 0.05;
    }
    /* rows: per chromosome count proportional to length; sequential placement with random gaps */
    S->n_rmsk = n_rmsk;
    S->rows = malloc(sizeof(srow_t) * (n_rmsk ? n_rmsk : 1));
    uint64_t made = 0;
    for (int c = 0; c < S->n_chrom; c++) {
        uint64_t want = (c == S->n_chrom - 1) ? n_rmsk - made
                        : (uint64_t)((double)n_rmsk * (double)S->lens[c] / (double)S->glen);
        if (made + want > n_rmsk) want = n_rmsk - made;
        S->chrom_row0[c] = made;
        rng_t q = rng_seed(seed, 0x7000 + c);
        uint32_t clen = S->lens[c];
        double mean_gap = want ? (double)clen / (double)want : 0;
        double pos = 0; uint32_t prev_s = 0, prev_e = 0;
        for (uint64_t k = 0; k < want; k++) {
            srow_t *w = &S->rows[made + k];
            uint32_t len;
            if (shape == 0) len = 30 + (uint32_t)rng_below(&q, 1471);
            else { double l = 200.0 * exp(0.9 * rng_norm(&q)); if (l < 12) l = 12; if (l > 160000) l = 160000; len = (uint32_t)l; }
            uint32_t s;
            if (k > 0 && rng_unif(&q) < 0.02 && prev_e > prev_s + 4) {
                /* nested / overlapping with the previous row */
                s = prev_s + (uint32_t)rng_below(&q, prev_e - prev_s);
            } else {
                pos += mean_gap * (0.1 + 1.8 * rng_unif(&q));
                s = (uint32_t)pos;
            }
            if (s < prev_s) s = prev_s;                 /* keep rows sorted by start */
            if (s >= clen - 12) s = clen - 12;
            uint32_t e = s + len; if (e > clen) e = clen;
            if (e <= s) e = s + 1;
            /* skewed subfamily usage: a few hot names take most rows (Alu/L1-like) */
            double u = rng_unif(&q);
            uint32_t sf = (uint32_t)((double)n_subfam * u * u * u);
            if (sf >= (uint32_t)n_subfam) sf = n_subfam - 1;
            uint32_t fam = sf % n_fam, cla = fam % n_cla;
            if (rng_unif(&q) < 0.005) { fam = (fam + 1) % n_fam; cla = (cla + 1) % n_cla; }
            uint32_t L = S->conslen[sf], el = e - s;
            uint32_t cs = (L > el) ? (uint32_t)rng_below(&q, L - el + 1) : 0;
            uint32_t ce = cs + el; if (ce > L) ce = L;
            w->chrom = c; w->start = s; w->end = e; w->subfam = sf; w->fam = fam; w->cla = cla;
            w->cs = cs; w->ce = ce; w->conslen = L; w->strand = (rng_next(&q) & 1) ? '+' : '-';
            prev_s = s; prev_e = e;
        }
        made += want;
    }
    S->chrom_row0[S->n_chrom] = made;
    S->n_rmsk = made;
    qsort(S->rows, S->n_rmsk, sizeof(srow_t), cmp_row);   /* already sorted; keeps the contract explicit */
    /* per-subfamily lists */
    S->sf_off = calloc(n_subfam + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < S->n_rmsk; i++) S->sf_off[S->rows[i].subfam + 1]++;
    for (int i = 0; i < n_subfam; i++) S->sf_off[i + 1] += S->sf_off[i];
    S->sf_rows = malloc(sizeof(uint64_t) * (S->n_rmsk ? S->n_rmsk : 1));
    uint64_t *fill = calloc(n_subfam, sizeof(uint64_t));
    for (uint64_t i = 0; i < S->n_rmsk; i++) { uint32_t sf = S->rows[i].subfam; S->sf_rows[S->sf_off[sf] + fill[sf]++] = i; }
    free(fill);
    return S;
}

void synth_free(void *h) {
    synth_t *S = h; if (!S) return;
    free(S->conslen); free(S->in_sizefile); free(S->rows); free(S->sf_off); free(S->sf_rows); free(S);
}

uint64_t synth_n_rmsk(void *h) { return ((synth_t *)h)->n_rmsk; }

int synth_write_sizes(void *h, const char *chrom_path, const char *rep_path) {
    synth_t *S = h;
    FILE *f = fopen(chrom_path, "w"); if (!f) return -1;
    for (int i = 0; i < S->n_chrom; i++) fprintf(f, "%s\t%u\n", S->names[i], S->lens[i]);
    fclose(f);
    f = fopen(rep_path, "w"); if (!f) return -1;
    for (int i = 0; i < S->n_subfam; i++) if (S->in_sizefile[i]) fprintf(f, "SUB%04d\t%u\n", i, S->conslen[i]);
    fclose(f);
    return 0;
}

int synth_write_rmsk(void *h, const char *path) {
    synth_t *S = h;
    FILE *f = fopen(path, "w"); if (!f) return -1;
    static char big[1 << 20]; setvbuf(f, big, _IOFBF, sizeof big);
    for (uint64_t i = 0; i < S->n_rmsk; i++) {
        srow_t *w = &S->rows[i];
        uint32_t L = w->conslen; long left = -(long)(L - w->ce);
        /* bin swScore milliDiv milliDel milliIns genoName genoStart genoEnd genoLeft strand repName repClass repFamily repStart repEnd repLeft id */
        if (w->strand == '+')
            fprintf(f, "585\t%u\t%u\t%u\t%u\t%s\t%u\t%u\t-%u\t+\tSUB%04u\tCLS%02u\tFAM%02u\t%u\t%u\t%ld\t%u\n",
                    300 + (uint32_t)(i % 4000), (uint32_t)(i % 300), (uint32_t)(i % 40), (uint32_t)(i % 30),
                    S->names[w->chrom], w->start, w->end, S->lens[w->chrom] - w->end,
                    w->subfam, w->cla, w->fam, w->cs, w->ce, left, (uint32_t)(i % 9 + 1));
        else
            fprintf(f, "585\t%u\t%u\t%u\t%u\t%s\t%u\t%u\t-%u\t-\tSUB%04u\tCLS%02u\tFAM%02u\t%ld\t%u\t%u\t%u\n",
                    300 + (uint32_t)(i % 4000), (uint32_t)(i % 300), (uint32_t)(i % 40), (uint32_t)(i % 30),
                    S->names[w->chrom], w->start, w->end, S->lens[w->chrom] - w->end,
                    w->subfam, w->cla, w->fam, left, w->ce, w->cs, (uint32_t)(i % 9 + 1));
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ BAM header */
uint64_t synth_bam_header(void *h, uint8_t *buf, uint64_t cap) {
    synth_t *S = h;
    char text[8192]; int tl = 0;
    tl += snprintf(text + tl, sizeof(text) - tl, "@HD\tVN:1.0\tSO:coordinate\n");
    for (int i = 0; i < S->n_chrom; i++) tl += snprintf(text + tl, sizeof(text) - tl, "@SQ\tSN:%s\tLN:%u\n", S->names[i], S->lens[i]);
    tl += snprintf(text + tl, sizeof(text) - tl, "@SQ\tSN:%s\tLN:%u\n", ABSENT_NAME, ABSENT_LEN);
    uint64_t need = 4 + 4 + tl + 4;
    for (int i = 0; i < S->n_chrom; i++) need += 4 + strlen(S->names[i]) + 1 + 4;
    need += 4 + strlen(ABSENT_NAME) + 1 + 4;
    if (!buf || cap < need) return need;
    uint8_t *p = buf;
    memcpy(p, "BAM\1", 4); p += 4;
    int32_t v = tl; memcpy(p, &v, 4); p += 4; memcpy(p, text, tl); p += tl;
    v = S->n_chrom + 1; memcpy(p, &v, 4); p += 4;
    for (int i = 0; i <= S->n_chrom; i++) {
        const char *nm = i < S->n_chrom ? S->names[i] : ABSENT_NAME;
        uint32_t ln = i < S->n_chrom ? S->lens[i] : ABSENT_LEN;
        v = (int32_t)strlen(nm) + 1; memcpy(p, &v, 4); p += 4; memcpy(p, nm, v); p += v;
        memcpy(p, &ln, 4); p += 4;
    }
    return (uint64_t)(p - buf);
}

/* ------------------------------------------------------------------ BAM records */
#define CHUNK_READS 65536ULL       /* reads (mode 0/1) or pairs (mode 2) per generation chunk */

static const uint8_t MAPQ_TAB[8] = {0, 0, 3, 20, 37, 37, 37, 60};

static inline int reg2bin(int beg, int end) {
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
}

/* emit one record; out == NULL only sizes it */
static size_t put_record(uint8_t *out, int32_t tid, int32_t pos, uint32_t mapq, uint32_t flag,
                         const char *qname, const uint32_t *cigar, int n_cigar, int l_seq,
                         int32_t mtid, int32_t mpos, int32_t isize,
                         const uint8_t *aux, int l_aux, rng_t *r) {
    int l_qname = (int)strlen(qname) + 1;
    size_t sz = 4 + 32 + l_qname + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq + l_aux;
    if (!out) {   /* sizing pass: advance the RNG exactly as the fill pass does */
        for (int i = 0; i < (l_seq + 1) / 2; i += 8) rng_next(r);
        for (int i = 0; i < l_seq; i += 8) rng_next(r);
        return sz;
    }
    uint32_t x[9];
    int end = pos;
    for (int k = 0; k < n_cigar; k++) { int op = cigar[k] & 0xf; if (op == 0 || op == 2 || op == 3) end += cigar[k] >> 4; }
    if (end == pos) end = pos + 1;
    int bin = pos >= 0 ? reg2bin(pos, end) : 4680;
    x[0] = (uint32_t)(sz - 4);
    x[1] = (uint32_t)tid; x[2] = (uint32_t)pos;
    x[3] = (uint32_t)bin << 16 | mapq << 8 | (uint32_t)l_qname;
    x[4] = flag << 16 | (uint32_t)n_cigar;
    x[5] = (uint32_t)l_seq; x[6] = (uint32_t)mtid; x[7] = (uint32_t)mpos; x[8] = (uint32_t)isize;
    memcpy(out, x, 36); uint8_t *p = out + 36;
    memcpy(p, qname, l_qname); p += l_qname;
    memcpy(p, cigar, 4 * (size_t)n_cigar); p += 4 * (size_t)n_cigar;
    int nb = (l_seq + 1) / 2;
    for (int i = 0; i < nb; i += 8) { uint64_t v = rng_next(r); int n = nb - i < 8 ? nb - i : 8;
        for (int j = 0; j < n; j++) { uint8_t b = (uint8_t)(v >> (8 * j)); p[i + j] = (uint8_t)((1u << (b & 3)) << 4 | (1u << ((b >> 2) & 3))); } }
    p += nb;
    for (int i = 0; i < l_seq; i += 8) { uint64_t v = rng_next(r); int n = l_seq - i < 8 ? l_seq - i : 8;
        for (int j = 0; j < n; j++) p[i + j] = (uint8_t)(2 + ((v >> (8 * j)) & 0x1f) + ((v >> (8 * j + 5)) & 7)); }
    p += l_seq;
    memcpy(p, aux, l_aux);
    return sz;
}

static inline void lin2chrom(const synth_t *S, uint64_t lin, int *c, uint32_t *pos) {
    int lo = 0, hi = S->n_chrom - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (S->cum[mid] <= lin) lo = mid; else hi = mid - 1; }
    *c = lo; *pos = (uint32_t)(lin - S->cum[lo]);
}

typedef struct { const synth_t *S; int mode; uint64_t n_units, seed; } bamcfg_t;

/* generate chunk `ci` (units [ci*CHUNK, min(n,(ci+1)*CHUNK)) ); out NULL => size only */
static size_t gen_chunk(const bamcfg_t *B, uint64_t ci, uint8_t *out, uint64_t *n_records) {
    const synth_t *S = B->S;
    uint64_t u0 = ci * CHUNK_READS, u1 = u0 + CHUNK_READS; if (u1 > B->n_units) u1 = B->n_units;
    uint64_t n = u1 - u0, nrec = 0;
    rng_t r = rng_seed(B->seed, 0x100000 + ci);
    /* linear range of this chunk: proportional slice of the genome -> coordinate-sorted stream */
    double g0 = (double)S->glen * ((double)u0 / (double)B->n_units);
    double g1 = (double)S->glen * ((double)u1 / (double)B->n_units);
    double step = n ? (g1 - g0) / (double)n : 0;
    size_t off = 0;
    char qn[32]; uint8_t aux[512]; uint32_t cig[4];
    int absent_tid = S->n_chrom;
    for (uint64_t k = 0; k < n; k++) {
        uint64_t id = u0 + k;
        uint64_t lin = (uint64_t)(g0 + ((double)k + rng_unif(&r)) * step);
        if (lin >= S->glen) lin = S->glen - 1;
        int c; uint32_t pos; lin2chrom(S, lin, &c, &pos);
        uint32_t mapq = MAPQ_TAB[rng_next(&r) & 7];
        double u = rng_unif(&r);
        if (B->mode == 0 || B->mode == 1) {
            int L = B->mode == 0 ? 50 : 75;
            snprintf(qn, sizeof qn, "r%09llu", (unsigned long long)id);
            uint32_t flag = (rng_next(&r) & 1) ? 16 : 0;
            int32_t tid = c, p = (int32_t)pos; int nc = 1; cig[0] = (uint32_t)L << 4;
            int la = 0;
            uint32_t nm = (uint32_t)rng_below(&r, 4);
            if (u < 0.02) {                       /* unmapped */
                flag = 4; tid = -1; p = -1; nc = 0; mapq = 0;
            } else if (u < 0.03) {                /* contig absent from the size file */
                tid = absent_tid; p = (int32_t)rng_below(&r, ABSENT_LEN - 200);
            } else if (u < 0.06) {                /* CIGARs that exercise bam_calend: D, N, S, I, =, X */
                int w = (int)(rng_next(&r) % 5);
                if (w == 0) { nc = 3; cig[0] = 20u << 4; cig[1] = (2u << 4) | 2; cig[2] = (uint32_t)(L - 20) << 4; }
                else if (w == 1) { nc = 2; cig[0] = (5u << 4) | 4; cig[1] = (uint32_t)(L - 5) << 4; }
                else if (w == 2) { nc = 3; cig[0] = 25u << 4; cig[1] = (1000u << 4) | 3; cig[2] = (uint32_t)(L - 25) << 4; }
                else if (w == 3) { nc = 3; cig[0] = 30u << 4; cig[1] = (3u << 4) | 1; cig[2] = (uint32_t)(L - 33) << 4; }
                else { nc = 2; cig[0] = (uint32_t)(L - 10) << 4 | 7; cig[1] = (10u << 4) | 8; }
            }
            if (B->mode == 1 && !(flag & 4) && tid != absent_tid && S->n_rmsk && u >= 0.06 && u < 0.36) {
                /* multi-read: MAPQ 0, primary inside a repeat, NM:i + XA:Z with 1-5 alternates */
                mapq = 0;
                uint64_t ri = S->chrom_row0[c] + rng_below(&r, S->chrom_row0[c + 1] - S->chrom_row0[c]);
                if (S->chrom_row0[c + 1] > S->chrom_row0[c]) {
                    const srow_t *w = &S->rows[ri];
                    uint32_t span = w->end - w->start;
                    p = (int32_t)(w->start + rng_below(&r, span));
                    if ((uint32_t)p + 300 >= S->lens[c]) p = (int32_t)w->start;
                    aux[la++] = 'N'; aux[la++] = 'M'; aux[la++] = 'i'; memcpy(aux + la, &nm, 4); la += 4;
                    if (rng_next(&r) & 1) { aux[la++] = 'X'; aux[la++] = '0'; aux[la++] = 'C'; aux[la++] = 1; }
                    aux[la++] = 'X'; aux[la++] = 'A'; aux[la++] = 'Z';
                    int nalt = 1 + (int)rng_below(&r, 5);
                    int same = (rng_next(&r) & 1);
                    for (int a = 0; a < nalt; a++) {
                        const srow_t *t;
                        if (same) { uint64_t m = S->sf_off[w->subfam + 1] - S->sf_off[w->subfam];
                                    t = &S->rows[S->sf_rows[S->sf_off[w->subfam] + rng_below(&r, m)]]; }
                        else t = &S->rows[rng_below(&r, S->n_rmsk)];
                        uint32_t ap = t->start + (uint32_t)rng_below(&r, t->end - t->start) + 1;
                        uint32_t anm = same ? nm : (uint32_t)rng_below(&r, 6);
                        la += snprintf((char *)aux + la, sizeof(aux) - la, "%s,%c%u,%dM,%u;", S->names[t->chrom],
                                       (rng_next(&r) & 1) ? '+' : '-', ap, L, anm);
                    }
                    aux[la++] = 0;
                }
            }
            if (la == 0) { aux[0] = 'N'; aux[1] = 'M'; aux[2] = 'C'; aux[3] = (uint8_t)nm; la = 4; }
            off += put_record(out ? out + off : NULL, tid, p, mapq, flag, qn, cig, nc, L, -1, -1, 0, aux, la, &r);
            nrec++;
        } else {
            /* paired-end 100: the two mates are emitted next to each other, leftmost first */
            int L = 100;
            snprintf(qn, sizeof qn, "p%010llu", (unsigned long long)id);
            aux[0] = 'N'; aux[1] = 'M'; aux[2] = 'C'; aux[3] = (uint8_t)rng_below(&r, 4);
            cig[0] = (uint32_t)L << 4;
            int fwd_first = (int)(rng_next(&r) & 1);     /* 99/147 or 83/163 */
            uint32_t mq2 = MAPQ_TAB[rng_next(&r) & 7];
            if (u < 0.01) {                        /* both unmapped: 77 / 141 */
                off += put_record(out ? out + off : NULL, -1, -1, 0, 77, qn, cig, 0, L, -1, -1, 0, aux, 4, &r);
                off += put_record(out ? out + off : NULL, -1, -1, 0, 141, qn, cig, 0, L, -1, -1, 0, aux, 4, &r);
            } else if (u < 0.03) {                 /* mate unmapped: 73 (or 89) + 133 placed at the mate's position */
                uint32_t f1 = 73 | ((rng_next(&r) & 1) ? 16 : 0);
                off += put_record(out ? out + off : NULL, c, (int32_t)pos, mapq, f1, qn, cig, 1, L, c, (int32_t)pos, 0, aux, 4, &r);
                off += put_record(out ? out + off : NULL, c, (int32_t)pos, 0, 133, qn, cig, 0, L, c, (int32_t)pos, 0, aux, 4, &r);
            } else {
                double is = 300.0 + 60.0 * rng_norm(&r);
                if (is < 101) is = 101;
                if (is > 499) is = 499;
                if (u < 0.06) is = 501 + (double)rng_below(&r, 3000);     /* over the -I threshold */
                int32_t isz = (int32_t)is;
                int32_t p1 = (int32_t)pos, p2 = p1 + isz - L;
                if ((uint32_t)p2 + L >= S->lens[c]) { p1 = (int32_t)pos - isz; if (p1 < 0) p1 = 0; p2 = p1 + isz - L; }
                if (fwd_first) {   /* read1 forward at p1 (flag 99), read2 reverse at p2 (147) */
                    off += put_record(out ? out + off : NULL, c, p1, mapq, 99, qn, cig, 1, L, c, p2, isz, aux, 4, &r);
                    off += put_record(out ? out + off : NULL, c, p2, mq2, 147, qn, cig, 1, L, c, p1, -isz, aux, 4, &r);
                } else {           /* read2 forward at p1 (163), read1 reverse at p2 (83) */
                    off += put_record(out ? out + off : NULL, c, p1, mq2, 163, qn, cig, 1, L, c, p2, isz, aux, 4, &r);
                    off += put_record(out ? out + off : NULL, c, p2, mapq, 83, qn, cig, 1, L, c, p1, -isz, aux, 4, &r);
                }
            }
            nrec += 2;
        }
    }
    if (n_records) *n_records = nrec;
    return off;
}

typedef struct { const bamcfg_t *B; uint64_t c0, c1; uint64_t *sizes, *nrec, *offs; uint8_t *buf; int tid, nth; int fill; } gjob_t;
static void *gen_worker(void *a) {
    gjob_t *J = a;
    for (uint64_t ci = J->c0 + J->tid; ci < J->c1; ci += J->nth) {
        if (!J->fill) J->sizes[ci - J->c0] = gen_chunk(J->B, ci, NULL, &J->nrec[ci - J->c0]);
        else gen_chunk(J->B, ci, J->buf + J->offs[ci - J->c0], NULL);
    }
    return NULL;
}
static void run_jobs(gjob_t *proto, int nth) {
    pthread_t th[256]; gjob_t jobs[256];
    if (nth > 256) nth = 256;
    if (nth < 1) nth = 1;
    for (int t = 0; t < nth; t++) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].nth = nth; pthread_create(&th[t], NULL, gen_worker, &jobs[t]); }
    for (int t = 0; t < nth; t++) pthread_join(th[t], NULL);
}

uint64_t synth_n_chunks(uint64_t n_units) { return (n_units + CHUNK_READS - 1) / CHUNK_READS; }

/* Size (bytes) and record count of chunks [c0,c1) of the stream of n_units reads (mode 0/1) or pairs (mode 2). */
uint64_t synth_records_size(void *h, int mode, uint64_t n_units, uint64_t seed, uint64_t c0, uint64_t c1, int nth, uint64_t *n_records) {
    bamcfg_t B = {h, mode, n_units, seed};
    uint64_t nc = c1 - c0; if (!nc) { if (n_records) *n_records = 0; return 0; }
    uint64_t *sizes = calloc(nc, 8), *nrec = calloc(nc, 8);
    gjob_t J = {&B, c0, c1, sizes, nrec, NULL, NULL, 0, 0, 0};
    run_jobs(&J, nth);
    uint64_t tot = 0, nr = 0; for (uint64_t i = 0; i < nc; i++) { tot += sizes[i]; nr += nrec[i]; }
    free(sizes); free(nrec);
    if (n_records) *n_records = nr;
    return tot;
}

/* Fill buf (capacity from synth_records_size) with chunks [c0,c1). Returns bytes written. */
uint64_t synth_records_fill(void *h, int mode, uint64_t n_units, uint64_t seed, uint64_t c0, uint64_t c1, uint8_t *buf, int nth) {
    bamcfg_t B = {h, mode, n_units, seed};
    uint64_t nc = c1 - c0; if (!nc) return 0;
    uint64_t *sizes = calloc(nc, 8), *nrec = calloc(nc, 8), *offs = calloc(nc + 1, 8);
    gjob_t J = {&B, c0, c1, sizes, nrec, offs, buf, 0, 0, 0};
    run_jobs(&J, nth);
    for (uint64_t i = 0; i < nc; i++) offs[i + 1] = offs[i] + sizes[i];
    J.fill = 1; run_jobs(&J, nth);
    uint64_t tot = offs[nc];
    free(sizes); free(nrec); free(offs);
    return tot;
}

/* ------------------------------------------------------------------ BGZF writer */
#define BGZF_PAYLOAD 0xff00u
typedef struct { const uint8_t *src; uint64_t n; uint64_t nblk; int level; uint8_t *dst; uint32_t *clen; int tid, nth; } zjob_t;
static size_t bgzf_block(uint8_t *dst, const uint8_t *src, uint32_t n, int level) {
    /* 18-byte header, raw deflate, CRC32 + ISIZE */
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(dst, hdr, 16);
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)src; zs.avail_in = n; zs.next_out = dst + 18; zs.avail_out = 65536 - 18 - 8;
    int st = deflate(&zs, Z_FINISH);
    if (st != Z_STREAM_END) {            /* incompressible: fall back to a stored block */
        deflateEnd(&zs); memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, 0, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = (Bytef *)src; zs.avail_in = n; zs.next_out = dst + 18; zs.avail_out = 65536 - 18 - 8;
        deflate(&zs, Z_FINISH);
    }
    size_t cl = zs.total_out; deflateEnd(&zs);
    uint32_t crc = (uint32_t)crc32(crc32(0, NULL, 0), src, n);
    size_t tot = 18 + cl + 8;
    uint16_t bs = (uint16_t)(tot - 1); memcpy(dst + 16, &bs, 2);
    memcpy(dst + 18 + cl, &crc, 4); memcpy(dst + 18 + cl + 4, &n, 4);
    return tot;
}
static void *z_worker(void *a) {
    zjob_t *J = a;
    for (uint64_t b = J->tid; b < J->nblk; b += J->nth) {
        uint64_t o = b * BGZF_PAYLOAD; uint32_t n = (uint32_t)((J->n - o) < BGZF_PAYLOAD ? (J->n - o) : BGZF_PAYLOAD);
        J->clen[b] = (uint32_t)bgzf_block(J->dst + b * 65536ULL, J->src + o, n, J->level);
    }
    return NULL;
}

/* Write header+records as one BGZF file. Blocks are cut every 0xff00 bytes of the concatenated
 * stream (records straddle blocks, as in real BAMs). level 0..9 (0 = stored). */
int synth_write_bam(const char *path, const uint8_t *hdr, uint64_t hdr_len, const uint8_t *recs, uint64_t rec_len, int level, int nth) {
    FILE *f = fopen(path, "wb"); if (!f) return -1;
    static const uint8_t eof_blk[28] = {31,139,8,4,0,0,0,0,0,0xff,6,0,'B','C',2,0,27,0,3,0,0,0,0,0,0,0,0,0};
    /* process in windows so the staging buffer stays bounded */
    const uint64_t WIN_BLK = 4096;
    uint8_t *stage = malloc(WIN_BLK * 65536ULL); uint32_t *clen = malloc(WIN_BLK * 4);
    uint8_t *cat = malloc(WIN_BLK * BGZF_PAYLOAD);
    uint64_t total = hdr_len + rec_len, done = 0;
    if (nth < 1) nth = 1;
    if (nth > 256) nth = 256;
    while (done < total) {
        uint64_t n = total - done; if (n > WIN_BLK * BGZF_PAYLOAD) n = WIN_BLK * BGZF_PAYLOAD;
        /* gather the window (header bytes first, then records) */
        uint64_t w = 0;
        if (done < hdr_len) { uint64_t k = hdr_len - done; if (k > n) k = n; memcpy(cat, hdr + done, k); w = k; }
        if (w < n) memcpy(cat + w, recs + (done + w - hdr_len), n - w);
        uint64_t nblk = (n + BGZF_PAYLOAD - 1) / BGZF_PAYLOAD;
        pthread_t th[256]; zjob_t jobs[256];
        for (int t = 0; t < nth; t++) { zjob_t j = {cat, n, nblk, level, stage, clen, t, nth}; jobs[t] = j; pthread_create(&th[t], NULL, z_worker, &jobs[t]); }
        for (int t = 0; t < nth; t++) pthread_join(th[t], NULL);
        for (uint64_t b = 0; b < nblk; b++) if (fwrite(stage + b * 65536ULL, 1, clen[b], f) != clen[b]) { fclose(f); return -2; }
        done += n;
    }
    fwrite(eof_blk, 1, 28, f);
    free(stage); free(clen); free(cat);
    return fclose(f) ? -3 : 0;
}

/* ------------------------------------------------------------------ CpG bedGraph (cfg4) */
int synth_write_bedgraph(void *h, const char *path, uint64_t n_rows, uint64_t seed) {
    synth_t *S = h;
    FILE *f = fopen(path, "w"); if (!f) return -1;
    static char big[1 << 20]; setvbuf(f, big, _IOFBF, sizeof big);
    rng_t r = rng_seed(seed, 0xc96);
    double step = (double)S->glen / (double)n_rows;
    for (uint64_t k = 0; k < n_rows; k++) {
        uint64_t lin = (uint64_t)(((double)k + rng_unif(&r)) * step); if (lin >= S->glen) lin = S->glen - 1;
        int c; uint32_t pos; lin2chrom(S, lin, &c, &pos);
        if (pos + 2 > S->lens[c]) pos = S->lens[c] - 2;
        uint32_t cents = (uint32_t)rng_below(&r, 5001);
        if (k % 1000 == 999) fprintf(f, "%s\t%u\t%u\t%u.%02u\n", ABSENT_NAME, pos % 100000, pos % 100000 + 2, cents / 100, cents % 100);
        else fprintf(f, "%s\t%u\t%u\t%u.%02u\n", S->names[c], pos, pos + 2, cents / 100, cents % 100);
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ CLI */
#ifdef ITX_SYNTH_MAIN
int main(int argc, char **argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: itx_synth <outdir> <shape 0|1> <n_rmsk> <n_subfam> <mode 0|1|2> <n_units> <seed> [level=1] [threads=8] [n_cpg=0]\n");
        return 1;
    }
    const char *od = argv[1]; int shape = atoi(argv[2]); uint64_t n_rmsk = strtoull(argv[3], 0, 0);
    int n_subfam = atoi(argv[4]), mode = atoi(argv[5]); uint64_t n_units = strtoull(argv[6], 0, 0), seed = strtoull(argv[7], 0, 0);
    int level = argc > 8 ? atoi(argv[8]) : 1, nth = argc > 9 ? atoi(argv[9]) : 8;
    uint64_t n_cpg = argc > 10 ? strtoull(argv[10], 0, 0) : 0;
    int n_fam = shape ? 56 : 60, n_cla = shape ? 21 : 12;
    void *S = synth_new(shape, n_rmsk, n_subfam, n_fam, n_cla, seed);
    char a[4096], b[4096];
    snprintf(a, sizeof a, "%s/chrom.sizes", od); snprintf(b, sizeof b, "%s/rep.sizes", od); synth_write_sizes(S, a, b);
    snprintf(a, sizeof a, "%s/rmsk.txt", od); synth_write_rmsk(S, a);
    if (n_units) {
        uint8_t hdr[16384]; uint64_t hl = synth_bam_header(S, hdr, sizeof hdr);
        uint64_t nc = synth_n_chunks(n_units), nrec = 0;
        uint64_t sz = synth_records_size(S, mode, n_units, seed, 0, nc, nth, &nrec);
        uint8_t *buf = malloc(sz ? sz : 1); synth_records_fill(S, mode, n_units, seed, 0, nc, buf, nth);
        snprintf(a, sizeof a, "%s/reads.bam", od);
        if (synth_write_bam(a, hdr, hl, buf, sz, level, nth)) { fprintf(stderr, "write failed\n"); return 2; }
        fprintf(stderr, "itx_synth: %llu records, %llu uncompressed bytes -> %s\n", (unsigned long long)nrec, (unsigned long long)sz, a);
        free(buf);
    }
    if (n_cpg) { snprintf(a, sizeof a, "%s/cpg.bedGraph", od); synth_write_bedgraph(S, a, n_cpg, seed); }
    synth_free(S);
    return 0;
}
#endif
