/* itx_sam.c -- SAM text input (-S): the text front end of the reference (samopen "r" -> sam_header_read + sam_read1,
 * cussamtools/sam.c:39-65, bam_import.c:216-236 and 237-456) restated as a converter from SAM text to the
 * uncompressed BAM byte stream the device scans.  Nothing is decided here: every alignment line becomes the record
 * samtools would have built in memory (same core fields, same CIGAR words, same aux encoding -- integers take the
 * smallest type that holds them, A/c/C all become 'A'), and the stream goes through the same kernels as a .bam.
 * Host C.  Plain or gzip-compressed text, as gzopen reads both.
 */
#define _GNU_SOURCE
#include "itx_internal.h"
#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { uint8_t *p; uint64_t n, cap; } sbuf;
static void sb_need(sbuf *b, uint64_t more) {
    if (b->n + more <= b->cap) return;
    while (b->n + more > b->cap) b->cap = b->cap ? b->cap * 2 : (1u << 20);
    b->p = (uint8_t *)realloc(b->p, b->cap);
}
static void sb_put(sbuf *b, const void *src, uint64_t n) { sb_need(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void sb_u32(sbuf *b, uint32_t v) { sb_put(b, &v, 4); }
static void sb_u8(sbuf *b, uint8_t v) { sb_put(b, &v, 1); }

/* one text line of any length (without its line end); 0 at the end of the file */
static int next_line(gzFile g, char **line, size_t *cap, size_t *len) {
    size_t n = 0;
    for (;;) {
        if (*cap < n + 4096) { *cap = *cap ? *cap * 2 : 65536; *line = (char *)realloc(*line, *cap); }
        if (!gzgets(g, *line + n, (int)(*cap - n))) { if (n == 0) return 0; break; }
        n += strlen(*line + n);
        if (n && (*line)[n - 1] == '\n') break;
    }
    while (n && ((*line)[n - 1] == '\n' || (*line)[n - 1] == '\r')) n--;
    (*line)[n] = 0; *len = n;
    return 1;
}
static uint32_t flag_of_char(unsigned char c) {          /* the pre-1.0 letter flags (bam_import.c:43-60) */
    switch (c) {
        case '1': return 0x40; case '2': return 0x80; case 'P': return 0x2; case 'R': return 0x20; case 'U': return 0x8;
        case 'd': return 0x400; case 'f': return 0x200; case 'p': return 0x1; case 'r': return 0x10; case 's': return 0x100; case 'u': return 0x4;
    }
    return 0;
}
static const unsigned char NT16[256] = {
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 1, 2, 4, 8, 15,15,15,15, 15,15,15,15, 15, 0,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15, 15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15, 15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15};
static int reg2bin(uint32_t beg, uint32_t end) {
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (int)(beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (int)(beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (int)(beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (int)(beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (int)(beg >> 26);
    return 0;
}
static int32_t tid_of(const itx_strtab *names, const char *s) { return itx_strtab_find(names, s); }
static int b_elem_size(int t) { return (t == 'C' || t == 'c' || t == 'A') ? 1 : (t == 'S' || t == 's') ? 2 : (t == 'I' || t == 'i' || t == 'f') ? 4 : 0; }

#define FAIL(...) do { snprintf(err, ITX_ERRLEN, __VA_ARGS__); rc = ITX_EFORMAT; goto out; } while (0)

int itx_sam_to_bam_stream(const char *path, uint8_t **out_bam, uint64_t *out_len, char err[ITX_ERRLEN]) {
    int rc = ITX_OK;
    gzFile g = gzopen(path, "rb");
    if (!g) { snprintf(err, ITX_ERRLEN, "Error\n[sam file %s: %s]", path, strerror(errno)); return ITX_EIO; }
    gzbuffer(g, 1 << 20);
    char *line = NULL; size_t cap = 0, len = 0; long lineno = 0;
    sbuf text = {0}, recs = {0}, rec = {0};
    itx_strtab names; itx_strtab_init(&names); uint32_t *lens = NULL; size_t nlens = 0, clens = 0;
    int have = next_line(g, &line, &cap, &len);
    /* header: every leading line that starts with '@'; the targets are the @SQ lines' SN / LN in file order */
    while (have && line[0] == '@') {
        lineno++;
        sb_put(&text, line, len); sb_u8(&text, '\n');
        if (strncmp(line, "@SQ", 3) == 0 && (line[3] == '\t' || line[3] == 0)) {
            const char *sn = NULL, *ln = NULL; size_t snl = 0;
            for (char *q = line + 3; *q;) {
                if (*q == '\t') q++;
                char *e = strchr(q, '\t'); size_t fl = e ? (size_t)(e - q) : strlen(q);
                if (fl >= 3 && q[2] == ':') { if (q[0] == 'S' && q[1] == 'N' && !sn) { sn = q + 3; snl = fl - 3; } else if (q[0] == 'L' && q[1] == 'N' && !ln) ln = q + 3; }
                q += fl;
            }
            if (sn && ln) {
                char *nm = strndup(sn, snl);
                itx_strtab_add(&names, nm); free(nm);
                if (nlens == clens) { clens = clens ? clens * 2 : 256; lens = (uint32_t *)realloc(lens, clens * 4); }
                lens[nlens++] = (uint32_t)atoi(ln);
            }
        }
        have = next_line(g, &line, &cap, &len);
    }
    for (; have; have = next_line(g, &line, &cap, &len)) {
        lineno++;
        if (len == 0) continue;                                     /* empty lines are skipped (bam_import.c:250-253) */
        char *F[12]; int nf = 0; char *q = line;
        while (nf < 11) { F[nf++] = q; char *t = strchr(q, '\t'); if (!t) { q = NULL; break; } *t = 0; q = t + 1; }
        if (nf < 11) FAIL("truncated SAM line %ld of %s", lineno, path);
        char *aux = q;                                              /* the optional fields, still tab separated, or NULL */
        rec.n = 0;
        const size_t l_qname = strlen(F[0]) + 1;
        if (l_qname > 255) FAIL("read name longer than 254 characters at line %ld of %s", lineno, path);
        long flag; { char *s; flag = strtol(F[1], &s, 0); if (*s) { flag = 0; for (s = F[1]; *s; ++s) flag |= (long)flag_of_char((unsigned char)*s); } }
        int32_t tid = tid_of(&names, F[2]);
        if (tid < 0 && strcmp(F[2], "*") != 0) {
            if (names.n == 0) FAIL("[sam_read1] missing header? Abort!");
            fprintf(stderr, "[sam_read1] reference '%s' is recognized as '*'.\n", F[2]);
        }
        int32_t pos = isdigit((unsigned char)F[3][0]) ? atoi(F[3]) - 1 : -1;
        uint32_t mapq = isdigit((unsigned char)F[4][0]) ? (uint32_t)atoi(F[4]) & 0xff : 0;
        /* cigar */
        uint32_t n_cigar = 0, cig[65536 / 4]; int bin; uint32_t qlen_cigar = 0;
        if (F[5][0] != '*') {
            for (char *s = F[5]; *s; ++s) {
                if (isalpha((unsigned char)*s) || *s == '=') ++n_cigar;
                else if (!isdigit((unsigned char)*s)) FAIL("Parse error at line %ld: invalid CIGAR character", lineno);
            }
            if (n_cigar > 16383) FAIL("Parse error at line %ld: too many CIGAR operations", lineno);
            char *s = F[5], *t;
            uint32_t end = (uint32_t)pos;
            for (uint32_t i = 0; i < n_cigar; i++) {
                long x = strtol(s, &t, 10); int op = toupper((unsigned char)*t), code;
                switch (op) { case 'M': code = 0; break; case 'I': code = 1; break; case 'D': code = 2; break; case 'N': code = 3; break; case 'S': code = 4; break;
                              case 'H': code = 5; break; case 'P': code = 6; break; case '=': code = 7; break; case 'X': code = 8; break;
                              default: FAIL("Parse error at line %ld: invalid CIGAR operation", lineno); }
                s = t + 1;
                cig[i] = (uint32_t)x << 4 | (uint32_t)code;
                if (code == 0 || code == 2 || code == 3) end += (uint32_t)x;                      /* bam_calend */
                if (code == 0 || code == 1 || code == 4 || code == 7 || code == 8) qlen_cigar += (uint32_t)x;   /* bam_cigar2qlen */
            }
            if (*s) FAIL("Parse error at line %ld: unmatched CIGAR operation", lineno);
            bin = reg2bin((uint32_t)pos, end);
        } else {
            if (!(flag & 4)) { fprintf(stderr, "Parse warning at line %ld: mapped sequence without CIGAR\n", lineno); flag |= 4; }
            bin = reg2bin((uint32_t)pos, (uint32_t)pos + 1);
        }
        int32_t mtid = strcmp(F[6], "=") ? tid_of(&names, F[6]) : tid;
        int32_t mpos = isdigit((unsigned char)F[7][0]) ? atoi(F[7]) - 1 : -1;
        int32_t isize = (F[8][0] == '-' || isdigit((unsigned char)F[8][0])) ? atoi(F[8]) : 0;
        int32_t l_qseq = 0;
        if (strcmp(F[9], "*")) {
            l_qseq = (int32_t)strlen(F[9]);
            if (n_cigar && (uint32_t)l_qseq != qlen_cigar) FAIL("Parse error at line %ld: CIGAR and sequence length are inconsistent", lineno);
        }
        if (strcmp(F[10], "*") && (size_t)l_qseq != strlen(F[10])) FAIL("Parse error at line %ld: sequence and quality are inconsistent", lineno);
        /* record: block_size, core, qname, cigar, seq, qual, aux */
        sb_u32(&rec, 0);
        sb_u32(&rec, (uint32_t)tid); sb_u32(&rec, (uint32_t)pos);
        sb_u32(&rec, (uint32_t)(bin & 0xffff) << 16 | mapq << 8 | (uint32_t)l_qname);
        sb_u32(&rec, ((uint32_t)flag & 0xffff) << 16 | n_cigar);
        sb_u32(&rec, (uint32_t)l_qseq); sb_u32(&rec, (uint32_t)mtid); sb_u32(&rec, (uint32_t)mpos); sb_u32(&rec, (uint32_t)isize);
        sb_put(&rec, F[0], l_qname);
        sb_put(&rec, cig, 4ull * n_cigar);
        {
            sb_need(&rec, (uint64_t)l_qseq + (uint64_t)(l_qseq + 1) / 2);
            uint8_t *p = rec.p + rec.n; memset(p, 0, (size_t)(l_qseq + 1) / 2);
            for (int32_t i = 0; i < l_qseq; i++) p[i / 2] |= (uint8_t)(NT16[(unsigned char)F[9][i]] << 4 * (1 - i % 2));
            p += (l_qseq + 1) / 2;
            if (strcmp(F[10], "*") == 0) memset(p, 0xff, (size_t)l_qseq); else for (int32_t i = 0; i < l_qseq; i++) p[i] = (uint8_t)(F[10][i] - 33);
            rec.n += (uint64_t)l_qseq + (uint64_t)(l_qseq + 1) / 2;
        }
        while (aux) {
            char *t = strchr(aux, '\t'); if (t) *t = 0;
            const size_t al = strlen(aux);
            if (al < 6 || aux[2] != ':' || aux[4] != ':') FAIL("Parse error at line %ld: missing colon in auxiliary data", lineno);
            const int type = aux[3]; const char *v = aux + 5;
            sb_put(&rec, aux, 2);
            if (type == 'A' || type == 'a' || type == 'c' || type == 'C') { sb_u8(&rec, 'A'); sb_u8(&rec, (uint8_t)v[0]); }
            else if (type == 'I' || type == 'i') {
                long long x = atoll(v);
                if (x < 0) {
                    if (x >= -127) { sb_u8(&rec, 'c'); int8_t y = (int8_t)x; sb_put(&rec, &y, 1); }
                    else if (x >= -32767) { sb_u8(&rec, 's'); int16_t y = (int16_t)x; sb_put(&rec, &y, 2); }
                    else { sb_u8(&rec, 'i'); int32_t y = (int32_t)x; sb_put(&rec, &y, 4); }
                } else {
                    if (x <= 255) { sb_u8(&rec, 'C'); sb_u8(&rec, (uint8_t)x); }
                    else if (x <= 65535) { sb_u8(&rec, 'S'); uint16_t y = (uint16_t)x; sb_put(&rec, &y, 2); }
                    else { sb_u8(&rec, 'I'); uint32_t y = (uint32_t)x; sb_put(&rec, &y, 4); }
                }
            } else if (type == 'f') { sb_u8(&rec, 'f'); float y = (float)atof(v); sb_put(&rec, &y, 4); }
            else if (type == 'd') { sb_u8(&rec, 'd'); double y = atof(v); sb_put(&rec, &y, 8); }
            else if (type == 'Z' || type == 'H') {
                if (type == 'H') {
                    if ((al - 5) % 2 == 1) FAIL("Parse error at line %ld: length of the hex string not even", lineno);
                    for (size_t i = 0; i < al - 5; i++) { int c = toupper((unsigned char)v[i]); if (!((c >= '0' && c <= '9') || (c >= 'A' && c <= 'F'))) FAIL("Parse error at line %ld: invalid hex character", lineno); }
                }
                sb_u8(&rec, (uint8_t)type); sb_put(&rec, v, al - 5); sb_u8(&rec, 0);
            } else if (type == 'B') {
                if (al < 8) FAIL("Parse error at line %ld: too few values in aux type B", lineno);
                const int st = v[0], es = b_elem_size(st); int32_t n = 0;
                for (const char *p = v + 1; *p; ++p) if (*p == ',') ++n;
                if (!es || st == 'A') FAIL("Parse error at line %ld: unrecognized array type", lineno);
                sb_u8(&rec, 'B'); sb_u8(&rec, (uint8_t)st); sb_put(&rec, &n, 4);
                char *p = (char *)v + 2; const char *vend = aux + al;
                while (p < vend) {
                    if (st == 'f') { float y = (float)strtod(p, &p); sb_put(&rec, &y, 4); }
                    else { long y = strtol(p, &p, 0); if (es == 1) { uint8_t b = (uint8_t)y; sb_put(&rec, &b, 1); } else if (es == 2) { uint16_t b = (uint16_t)y; sb_put(&rec, &b, 2); } else { uint32_t b = (uint32_t)y; sb_put(&rec, &b, 4); } }
                    ++p;
                }
            } else FAIL("Parse error at line %ld: unrecognized type", lineno);
            aux = t ? t + 1 : NULL;
        }
        { uint32_t bs = (uint32_t)(rec.n - 4); memcpy(rec.p, &bs, 4); }
        sb_put(&recs, rec.p, rec.n);
    }
    {   /* magic, l_text, text, n_ref, {l_name, name, l_ref} */
        sbuf o = {0};
        sb_put(&o, "BAM\1", 4); sb_u32(&o, (uint32_t)text.n); sb_put(&o, text.p, text.n); sb_u32(&o, (uint32_t)names.n);
        for (int32_t i = 0; i < names.n; i++) { uint32_t l = (uint32_t)strlen(names.names[i]) + 1; sb_u32(&o, l); sb_put(&o, names.names[i], l); sb_u32(&o, lens[i]); }
        sb_put(&o, recs.p, recs.n);
        sb_need(&o, 64); memset(o.p + o.n, 0, 64);                   /* readers of the stream may look a few bytes past its end */
        *out_bam = o.p; *out_len = o.n;
    }
out:
    gzclose(g); free(line); free(text.p); free(recs.p); free(rec.p); free(lens); itx_strtab_free(&names);
    return rc;
}

int itx_sam_to_bam(const char *sam, uint8_t **bam, uint64_t *len, char err[ITX_ERRLEN]) {
    char lerr[ITX_ERRLEN]; if (!err) err = lerr;
    err[0] = 0;
    return itx_sam_to_bam_stream(sam, bam, len, err);
}
void itx_free(void *p) { free(p); }
