/* iteres_oracle.h -- CPU restatement of the iteres hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (iteres_b200/csrc) never links, loads or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors.  This
 * restatement is pinned against the UNMODIFIED reference binary (oracle/_ref/iteres, built by
 * oracle/Makefile from /root/reference) on the known-answer inputs of SURVEY.md 4.2 and on
 * generated inputs: tests/test_oracle_vs_reference.py compares every output file byte for byte,
 * and tests/golden/ holds the reference's outputs for the known-answer inputs so the same check
 * runs where /root/reference does not exist.
 */
#ifndef ITERES_ORACLE_H
#define ITERES_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_index ora_index;

/* Same knobs, same meaning and defaults as samFiles2nodupRepbedFileNew's arguments (generic.c:700). */
typedef struct {
    uint32_t mapQ;            /* -Q, default 10 */
    int32_t  filter;          /* 0 = stat (subfamily/family/class counters), 1 = filter (per-element) */
    int32_t  rmDup;           /* -R */
    int32_t  addChr;          /* -C */
    int32_t  discardWrongEnd; /* -D */
    uint32_t iSize;           /* -I, default 500 */
    uint32_t extension;       /* -E, default 150 */
    float    minCoverage;     /* -c, default 1e-4f */
    int32_t  treat;           /* -T */
    int32_t  diffSubfam;      /* !-x, default 1 for stat, 0 for filter */
} ora_opts;

/* Per-record trace (optional), used to check the decode and overlap kernels separately. */
#define ORA_T_FRAGMENT 1u     /* record produced a fragment that went to the overlap stage */
#define ORA_T_UNIQ     2u
#define ORA_T_MINUS    4u
#define ORA_T_HAS_XA   8u
#define ORA_T_DIFFSUB  16u    /* discarded by mapped2diffSubfam */
#define ORA_T_COUNTED  32u    /* counted as a read in repeats */
typedef struct {
    uint32_t start, end;      /* fragment */
    int32_t  tid;             /* BAM reference id of the record */
    int32_t  sel_row;         /* rmsk row (0-based index among parsed rows) selected, or -1 */
    uint32_t flags;
} ora_trace;

/* rmsk2binKeeperHash (generic.c:1578-1707) + hashNameIntFile (obscure.c:139-150). NULL + err on failure. */
ora_index *ora_index_build(const char *chrom_sizes, const char *rep_sizes, const char *rmsk,
                           int filter_field, const char *filter_name, char err[256]);
void ora_index_free(ora_index *ix);
void ora_index_reset_counts(ora_index *ix);

/* samFiles2nodupRepbedFileNew over an UNCOMPRESSED BAM byte stream (magic+header+records).
 * Counters and the nochr/dup sets persist across calls until ora_index_reset_counts (multi-file runs).
 * trace may be NULL; at most trace_cap entries are written; *n_records gets the record count. */
int ora_scan_bam_stream(ora_index *ix, const uint8_t *bam, uint64_t len, const ora_opts *o,
                        uint64_t cnt[13], ora_trace *trace, uint64_t trace_cap, uint64_t *n_records);
/* Same, reading a BGZF-compressed .bam with zlib the way bgzf.c:471-565 does. */
int ora_scan_bam_file(ora_index *ix, const char *path, const ora_opts *o, uint64_t cnt[13]);
/* Inflate a whole .bam into a malloc'd buffer (caller frees with ora_free). */
uint8_t *ora_inflate_bam(const char *path, uint64_t *len);
void ora_free(void *p);

/* cpgBedGraphOverlapRepeat (generic.c:1064-1139). */
int ora_scan_cpg(ora_index *ix, const char *bedgraph, int filter, uint32_t *cpg_lines, uint32_t *cpg_in_repeat, char err[256]);

/* Writers: byte-identical restatements of writeWigandStat / writeReport / writeFilterOut /
 * MREwriteWigandStat / writeFilterOutMRE (generic.c:53-152, 1709-1771). */
int ora_write_stat(ora_index *ix, const char *subfam_stat, const char *wig, const char *fam_stat,
                   const char *class_stat, const char *wig_unique, uint64_t reads_num, uint64_t reads_num_unique);
int ora_write_report(const char *path, const uint64_t cnt[13], uint32_t mapQ, const char *subfam);
int ora_write_filter(ora_index *ix, const char *path, int readlist, int threshold, uint64_t reads_num);
int ora_write_cpg_stat(ora_index *ix, const char *subfam_stat, const char *wig, const char *fam_stat, const char *class_stat);
int ora_write_cpg_filter(ora_index *ix, const char *path, double score_threshold);

/* Single-query entry point for property tests of the overlap kernel: binKeeperFind order +
 * "last ascent" selection + minCoverage (generic.c:945-970).  chrom by name.  Returns the selected
 * rmsk row or -1; *n_hits gets the hit-list length; hits[] (cap entries) the rows in list order. */
int32_t ora_find_select(ora_index *ix, const char *chrom, uint32_t start, uint32_t end, float min_cov,
                        int32_t *n_hits, int32_t *hits, int32_t cap);

/* Plain accessors so tests can compare counters without parsing files. */
int32_t ora_n_subfam(ora_index *ix);
int32_t ora_n_fam(ora_index *ix);
int32_t ora_n_class(ora_index *ix);
int64_t ora_n_elem(ora_index *ix);
/* which: 0 subfamily, 1 family, 2 class; i in Kent hash-iteration order */
const char *ora_name(ora_index *ix, int which, int32_t i);
void ora_counts(ora_index *ix, int which, int32_t i, uint64_t out[4] /* read_count, unique, total_length, genome_count */);
uint32_t ora_subfam_length(ora_index *ix, int32_t i);
const uint32_t *ora_subfam_bp(ora_index *ix, int32_t i, int unique);

#ifdef __cplusplus
}
#endif
#endif
