/* itx_ordered.h -- the host pass behind the order-dependent side outputs: -B / -V bed lines (generic.c:925-936)
 * and the read names filter -r prints per locus (generic.c:662-666, 1726-1733).
 *
 * Nothing about a read is decided here: the device leaves one trace entry per record in file order (fragment,
 * strand, selected rmsk row, verdict flags); this walks the same records in a host copy of the stream and prints
 * their names and aux strings.  Plain host C++ shared by libiteres_gpu (itx_gpu.cu) and the test-only emulator.
 */
#ifndef ITX_ORDERED_H
#define ITX_ORDERED_H
#include "itx_logic.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

struct itx_ordered_sink {
    FILE *bed, *bed_u;          /* -B, -V (either may be NULL) */
    int names;                  /* filter -r */
    uint32_t mapQ;
    char **tname; int32_t n_ref;
    itx_index *ix;
};
/* the chromosome name the reference prints: target_name[tid] after the -C renaming (generic.c:781-791) */
static inline char **itx_ordered_tnames(const itx_bam_header *h) {
    char **tn = (char **)calloc((size_t)(h->n_ref ? h->n_ref : 1), sizeof(char *));
    for (int32_t t = 0; t < h->n_ref; t++) {
        const char *raw = h->names[t]; char nm[600];
        if (h->addChr && strcasecmp(raw, "MT") == 0) snprintf(nm, sizeof nm, "chrM");
        else if (h->addChr && strncmp(raw, "chr", 3) != 0) snprintf(nm, sizeof nm, "chr%s", raw);
        else snprintf(nm, sizeof nm, "%s", raw);
        tn[t] = strdup(nm);
    }
    return tn;
}
static inline void itx_ordered_names_init(itx_index *ix) {
    if (ix->el_names) return;
    const size_t ne = (size_t)(ix->n_elem ? ix->n_elem : 1);
    ix->el_names = (char ***)calloc(ne, sizeof(char **)); ix->el_names_n = (uint32_t *)calloc(ne, 4); ix->el_names_cap = (uint32_t *)calloc(ne, 4);
}
/* buf[0, nbytes) holds n records back to back (plus >= 64 readable bytes of slack), tr their trace entries.
 * Returns the bytes walked. */
static inline uint64_t itx_ordered_walk(const itx_ordered_sink &K, const uint8_t *buf, uint64_t nbytes, const itx_trace *tr, uint64_t n) {
    const itx_src_global G{buf};
    itx_index *ix = K.ix;
    uint64_t p = 0;
    for (uint64_t r = 0; r < n; r++) {
        if (p + 36 > nbytes) break;
        uint32_t x[9]; G.core(p, x);
        const uint64_t rec_end = p + 4 + (uint64_t)x[0];
        if ((int32_t)x[0] < 32 || rec_end > nbytes) break;
        const itx_trace T = tr[r];
        const char *qname = (const char *)(buf + p + 36);
        const uint32_t mapq = (x[3] >> 8) & 0xff;
        if ((T.flags & ITX_T_FRAGMENT) && (K.bed || K.bed_u) && T.tid >= 0 && T.tid < K.n_ref) {
            const char strand = (T.flags & ITX_T_MINUS) ? '-' : '+';
            if (K.bed) {
                fprintf(K.bed, "%s\t%u\t%u\t%s\t%i\t%c", K.tname[T.tid], T.start, T.end, qname, (int)mapq, strand);
                uint64_t a0, aend; itx_aux_range(p, x, &a0, &aend);
                const uint64_t xa = itx_aux_find(G, a0, aend, 'X', 'A');
                if (xa) {
                    const int32_t nm = itx_aux2i(G, itx_aux_find(G, a0, aend, 'N', 'M'), aend);
                    const uint8_t ty = xa < aend ? buf[xa] : 0;
                    if (ty == 'Z' || ty == 'H') { fprintf(K.bed, "\t%i\t", nm); for (uint64_t q = xa + 1; q < aend && buf[q]; q++) fputc(buf[q], K.bed); }
                    else fprintf(K.bed, "\t%i\t(null)", nm);            /* bam_aux2Z gives NULL for other types; glibc prints it so */
                }
                fputc('\n', K.bed);
            }
            if (K.bed_u && mapq >= K.mapQ) fprintf(K.bed_u, "%s\t%u\t%u\t%s\t%i\t%c\n", K.tname[T.tid], T.start, T.end, qname, (int)mapq, strand);
        }
        if (K.names && (T.flags & ITX_T_COUNTED) && T.sel_row >= 0 && (long long)T.sel_row < ix->n_rows) {
            const long long el = ix->row2el[T.sel_row];
            if (el >= 0) {
                if (ix->el_names_n[el] == ix->el_names_cap[el]) {
                    ix->el_names_cap[el] = ix->el_names_cap[el] ? ix->el_names_cap[el] * 2 : 4;
                    ix->el_names[el] = (char **)realloc(ix->el_names[el], sizeof(char *) * ix->el_names_cap[el]);
                }
                ix->el_names[el][ix->el_names_n[el]++] = strdup(qname);
            }
        }
        p = rec_end;
    }
    return p;
}
#endif
