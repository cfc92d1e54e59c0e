/* itx_logic.cuh -- the per-record / per-query device functions of the iteres hot path.
 *
 * Everything here is `ITX_HD` (__host__ __device__) so that the very same code can also be compiled
 * by g++ into the test-only emulator under tests/emu/ (unit tests of the device logic on a box
 * without a GPU).  The product library (itx_gpu.cu) only ever calls these from kernels.
 *
 * What is restated (reference file:line, lidaof/iteres v0.3.3-r123):
 *   bam_read1 / core unpack   cussamtools/bam.c:179-210, bam.h:169-178
 *   bam_calend                cussamtools/bam.c:17-27            (M, D, N only)
 *   bam_aux_get / aux2i       cussamtools/bam_aux.c:28-48, 159-170, bam.h:754-760
 *   per-read fragment logic   generic.c:748-905
 *   binKeeperFind             cuskent/binRange.c:196-227         (order: level 5->0, bin high->low, row old->new)
 *   getCov + selection        generic.c:296-301, 950-970         ("last ascent")
 *   mapped2diffSubfam         generic.c:303-341
 *   accumulation              generic.c:983-1024
 */
#ifndef ITX_LOGIC_CUH
#define ITX_LOGIC_CUH
#include "itx_internal.h"

#if defined(__CUDACC__)
#define ITX_HD __host__ __device__ __forceinline__
#define ITX_HDM __host__ __device__ __forceinline__
#define ITX_HDN __host__ __device__ __noinline__
#else
#define ITX_HD static inline
#define ITX_HDM inline
#define ITX_HDN static
#endif

/* ------------------------------------------------------------------ byte sources */
ITX_HD uint32_t itx_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
/* The BAM stream in global memory, addressed by stream offset.  Unaligned reads may touch up to 15 bytes
 * past the last byte asked for (stream buffers carry 64 bytes of slack). */
struct itx_src_global {
    const uint8_t *b;
    ITX_HDM uint8_t u8(uint64_t off) const {
#if defined(__CUDA_ARCH__)
        return __ldg(b + off);
#else
        return b[off];
#endif
    }
    ITX_HDM uint32_t w32(uint64_t aligned_off) const {
#if defined(__CUDA_ARCH__)
        return __ldg(reinterpret_cast<const uint32_t *>(b + aligned_off));
#else
        return *reinterpret_cast<const uint32_t *>(b + aligned_off);
#endif
    }
    ITX_HDM uint32_t u32(uint64_t off) const {
        const uint64_t a = off & ~3ull; const uint32_t sh = (uint32_t)(off & 3) * 8;
        const uint32_t lo = w32(a);
        if (sh == 0) return lo;
        return itx_funnel_r(lo, w32(a + 4), sh);
    }
    /* block_size + the 32-byte core of the record at p as nine u32 (x0 block_size, x1 refID, x2 pos,
     * x3 bin<<16|mapq<<8|l_qname, x4 flag<<16|n_cigar, x5 l_seq, x6 mtid, x7 mpos, x8 isize): four 16-byte
     * loads of the enclosing aligned window, then a byte realignment in registers */
    ITX_HDM void core(uint64_t p, uint32_t x[9]) const {
        uint32_t w[16];
        const uint32_t o = (uint32_t)(p & 15);
#if defined(__CUDA_ARCH__)
        const uint4 *v = reinterpret_cast<const uint4 *>(b + (p & ~15ull));
        const uint4 a0 = __ldg(v), a1 = __ldg(v + 1), a2 = __ldg(v + 2);
        uint4 a3 = make_uint4(0, 0, 0, 0);
        if (o > 12) a3 = __ldg(v + 3);
        w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w; w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
        w[8] = a2.x; w[9] = a2.y; w[10] = a2.z; w[11] = a2.w; w[12] = a3.x; w[13] = a3.y; w[14] = a3.z; w[15] = a3.w;
#else
        const uint32_t *v = reinterpret_cast<const uint32_t *>(b + (p & ~15ull));
        for (int i = 0; i < 16; i++) w[i] = (i < 12 || o > 12) ? v[i] : 0;
#endif
        const uint32_t q = o >> 2, sh = (o & 3) * 8;
        uint32_t t[12];
#pragma unroll
        for (int i = 0; i < 12; i++) t[i] = itx_funnel_r(w[i], w[i + 1], sh);
#pragma unroll
        for (int j = 0; j < 9; j++) x[j] = q == 0 ? t[j] : (q == 1 ? t[j + 1] : (q == 2 ? t[j + 2] : t[j + 3]));
    }
};
/* A power-of-two ring (shared memory on the device) holding stream bytes at (offset & mask). */
struct itx_src_ring {
    const uint8_t *ring; uint32_t mask;
    ITX_HDM uint8_t u8(uint64_t off) const { return ring[(uint32_t)off & mask]; }
    ITX_HDM uint32_t w32(uint64_t aligned_off) const { return *reinterpret_cast<const uint32_t *>(ring + ((uint32_t)aligned_off & mask)); }
    ITX_HDM uint32_t u32(uint64_t off) const {
        const uint64_t a = off & ~3ull; const uint32_t sh = (uint32_t)(off & 3) * 8;
        return itx_funnel_r(w32(a), w32(a + 4), sh);
    }
    ITX_HDM void core(uint64_t p, uint32_t x[9]) const {
        const uint64_t a = p & ~3ull; const uint32_t sh = (uint32_t)(p & 3) * 8;
        uint32_t w[10];
#pragma unroll
        for (int i = 0; i < 10; i++) w[i] = w32(a + 4u * i);
#pragma unroll
        for (int j = 0; j < 9; j++) x[j] = itx_funnel_r(w[j], w[j + 1], sh);
    }
};

/* Bytes staged linearly (shared memory on the device) when the caller knows that everything it will read lies inside the staged
 * bytes: buf[0] is the byte at offset `base`; no bounds tests */
struct itx_src_flat {
    const uint8_t *buf; unsigned long long base;
    ITX_HDM uint8_t u8(uint64_t off) const { return buf[(uint32_t)off - (uint32_t)base]; }
    ITX_HDM uint32_t w32(uint64_t aligned_off) const { return *reinterpret_cast<const uint32_t *>(buf + ((uint32_t)aligned_off - (uint32_t)base)); }
    ITX_HDM uint32_t u32(uint64_t off) const {
        const uint64_t a = off & ~3ull; const uint32_t sh = (uint32_t)(off & 3) * 8;
        return itx_funnel_r(w32(a), w32(a + 4), sh);
    }
};

/* ------------------------------------------------------------------ record chain */
/* Cheap structural test used ONLY to guess where a chunk's first record starts; the guess is checked
 * against the real chain afterwards (k_verify / k_fixup), so a wrong answer costs time, not exactness. */
template <class Src>
ITX_HD bool itx_plausible(const Src &S, uint64_t p, uint64_t len, int32_t n_ref, uint64_t *next) {
    if (p + 36 > len) return false;
    uint32_t bs = S.u32(p);
    if (bs < 33u || bs > (1u << 26)) return false;
    if (p + 4 + (uint64_t)bs > len) return false;
    int32_t tid = (int32_t)S.u32(p + 4);
    if (tid < -1 || tid >= n_ref) return false;
    int32_t pos = (int32_t)S.u32(p + 8);
    if (pos < -1) return false;
    uint32_t lq = S.u32(p + 12) & 0xff;
    if (lq == 0) return false;
    uint32_t nc = S.u32(p + 16) & 0xffff;
    int32_t ls = (int32_t)S.u32(p + 20);
    if (ls < 0) return false;
    int32_t mtid = (int32_t)S.u32(p + 24);
    if (mtid < -1 || mtid >= n_ref) return false;
    int32_t mpos = (int32_t)S.u32(p + 28);
    if (mpos < -1) return false;
    uint64_t need = 32ull + lq + 4ull * nc + (uint64_t)((ls + 1) / 2) + (uint64_t)ls;
    if (need > bs) return false;
    if (S.u8(p + 36 + lq - 1) != 0) return false;          /* qname is NUL terminated */
    *next = p + 4 + bs;
    return true;
}
/* The same test on a core that is already in registers (x as S.core() gives it): no early exits, so a warp that tests 32
 * offsets at once stays converged; only the qname terminator is left to the caller (it needs one more byte: *lq_out). */
ITX_HD bool itx_plausible_core(const uint32_t x[9], uint64_t p, uint64_t len, int32_t n_ref, uint32_t *lq_out, uint64_t *next) {
    const uint32_t bs = x[0];
    const int32_t tid = (int32_t)x[1], pos = (int32_t)x[2], ls = (int32_t)x[5], mtid = (int32_t)x[6], mpos = (int32_t)x[7];
    const uint32_t lq = x[3] & 0xff, nc = x[4] & 0xffff;
    const uint64_t need = 32ull + lq + 4ull * nc + (uint64_t)((ls + 1) / 2) + (uint64_t)ls;
    bool ok = bs >= 33u && bs <= (1u << 26);
    ok = ok && p + 4 + (uint64_t)bs <= len;
    ok = ok && tid >= -1 && tid < n_ref && pos >= -1 && lq != 0 && ls >= 0 && mtid >= -1 && mtid < n_ref && mpos >= -1;
    ok = ok && need <= bs;
    *lq_out = lq; *next = p + 4 + bs;
    return ok;
}
/* itx_plausible2's verdict through the core-first test: what the span kernels ask about 32 offsets per step */
template <class Src>
ITX_HD bool itx_plausible2_core(const Src &S, const uint32_t x[9], uint64_t p, uint64_t len, int32_t n_ref) {
    uint32_t lq; uint64_t nx, nx2;
    bool ok = itx_plausible_core(x, p, len, n_ref, &lq, &nx);
    if (ok) ok = S.u8(p + 36 + lq - 1) == 0 && (nx == len || itx_plausible(S, nx, len, n_ref, &nx2));
    return ok;
}
/* The chain walk of a staged piece of the stream (k_scan / k_decode_span), one step, one lane.  `buf` holds the stage,
 * offsets are relative to it; the record at q has size sz0 (already validated), records after it are predicted to have
 * size szp.  Lane 0 stands for the record at q; lane k >= 1 looks where record k would start under the prediction and
 * reports whether a record of exactly the predicted size starts there (inside [0, qh) and ending within room32).  The run
 * of lanes 0..r-1 that report true is accepted by the caller: each accepted start is the previous start plus a verified
 * size, i.e. the exact chain. */
ITX_HD uint32_t itx_buf_u32(const uint8_t *buf, uint32_t q) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + (q & ~3u));
    return itx_funnel_r(w[0], w[1], (q & 3u) * 8u);
}
ITX_HD bool itx_chain_lane(const uint8_t *buf, uint32_t q, uint32_t sz0, uint32_t szp, uint32_t lane, uint32_t qh, uint32_t room32, uint32_t *pk_out) {
    const uint32_t pk = lane ? q + sz0 + (lane - 1u) * szp : q;
    bool same = true;
    if (lane) {
        same = false;
        if (pk < qh && pk + szp <= room32) same = itx_buf_u32(buf, pk) + 4u == szp;
    }
    *pk_out = pk;
    return same;
}
/* a record start followed by another one (or by the end of the stream) */
template <class Src>
ITX_HD bool itx_plausible2(const Src &S, uint64_t p, uint64_t len, int32_t n_ref) {
    uint64_t nx, nx2;
    if (!itx_plausible(S, p, len, n_ref, &nx)) return false;
    return nx == len || itx_plausible(S, nx, len, n_ref, &nx2);
}
/* first offset in [lo, hi) that passes itx_plausible2 */
template <class Src>
ITX_HD uint64_t itx_speculate_entry(const Src &S, uint64_t lo, uint64_t hi, uint64_t len, int32_t n_ref) {
    /* slide over aligned words: every word is loaded once, the candidate block_size at each of its four
     * byte offsets comes out of a funnel shift, and only values in range go through the full test */
    uint64_t a = lo & ~3ull;
    uint32_t w0 = S.w32(a);
    for (; a < hi; a += 4) {
        const uint32_t w1 = S.w32(a + 4);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint64_t p = a + k;
            const uint32_t bs = itx_funnel_r(w0, w1, 8u * k);
            if (bs - 33u <= (1u << 26) - 33u && p >= lo && p < hi && itx_plausible2(S, p, len, n_ref)) return p;
        }
        w0 = w1;
    }
    return ITX_OFF_NONE;
}

/* ------------------------------------------------------------------ aux fields */
/* bam_aux_get: returns the offset of the type byte of `tag`, or 0.  Faithful to the reference's skip
 * rule, including its quirk that the type is upper-cased before its size is looked up (so 'f' and
 * 'd' values are skipped as size 0).  Reads are bounded by aend. */
/* Off: the type offsets are kept in -- uint64_t stream offsets, or uint32_t offsets into a staged copy (k_xa: half the
 * instructions per byte of a 64-bit walk) */
template <class Src, class Off>
ITX_HD Off itx_aux_find(const Src &S, Off s, Off aend, uint8_t t0, uint8_t t1) {
    while (s + 1 < aend) {
        uint8_t c0 = S.u8(s), c1 = S.u8(s + 1);
        s += 2;
        if (c0 == t0 && c1 == t1) return s;
        if (s >= aend) break;
        uint8_t ty = S.u8(s);
        if (ty >= 'a' && ty <= 'z') ty = (uint8_t)(ty - 32);
        s += 1;
        if (ty == 'Z' || ty == 'H') { while (s < aend && S.u8(s) != 0) s++; s++; }
        else if (ty == 'B') {
            if (s + 5 > aend) break;
            uint8_t sub = S.u8(s);
            uint32_t sz = (sub == 'C' || sub == 'c' || sub == 'A') ? 1u : (sub == 'S' || sub == 's') ? 2u : (sub == 'I' || sub == 'i' || sub == 'f') ? 4u : 0u;
            uint32_t n = (uint32_t)S.u8(s + 1) | (uint32_t)S.u8(s + 2) << 8 | (uint32_t)S.u8(s + 3) << 16 | (uint32_t)S.u8(s + 4) << 24;
            const uint64_t step = 5 + (uint64_t)sz * (uint64_t)(int64_t)(int32_t)n;
            /* a step that leaves the aux area (a corrupt count) wraps differently in 32 bits: the caller repeats the walk with 64-bit offsets */
            if (sizeof(Off) < 8 && step > (uint64_t)(aend - s)) return (Off) ~(Off)0;
            s += (Off)step;
        } else s += (ty == 'C' || ty == 'A') ? 1u : (ty == 'S') ? 2u : (ty == 'I') ? 4u : 0u;
    }
    return 0;
}
/* bam_aux2i on the value whose type byte sits at offset s (0 -> 0) */
template <class Src, class Off>
ITX_HD int32_t itx_aux2i(const Src &S, Off s, Off aend) {
    if (!s || s >= aend) return 0;
    uint8_t ty = S.u8(s); s++;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) if (s + (Off)i < aend) v |= (uint32_t)S.u8(s + (Off)i) << (8 * i);
    if (ty == 'c') return (int32_t)(int8_t)(v & 0xff);
    if (ty == 'C') return (int32_t)(v & 0xff);
    if (ty == 's') return (int32_t)(int16_t)(v & 0xffff);
    if (ty == 'S') return (int32_t)(v & 0xffff);
    if (ty == 'i' || ty == 'I') return (int32_t)v;
    return 0;
}
/* offsets of the aux area of the record at p (x = its core) */
ITX_HD void itx_aux_range(uint64_t p, const uint32_t x[9], uint64_t *a0, uint64_t *aend) {
    uint32_t lq = x[3] & 0xff, nc = x[4] & 0xffff; int32_t ls = (int32_t)x[5];
    *aend = p + 4 + (uint64_t)x[0];
    int64_t a = (int64_t)p + 36 + lq + 4ll * nc + (int64_t)ls + (int64_t)((ls + 1) / 2);
    *a0 = (a < (int64_t)p + 36 || (uint64_t)a > *aend) ? *aend : (uint64_t)a;
}

/* ------------------------------------------------------------------ one BAM record -> tuple */
ITX_HD uint32_t itx_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* WANT_XA = false: the caller looks for the XA tag itself (k_scan), whatever o.diffSubfam says */
template <class Src, bool WANT_XA = true>
ITX_HD itx_tuple itx_decode_record(const Src &S, uint64_t p, const uint32_t x[9], uint32_t rec_off,
                                   const itx_tidinfo *tidtab, int32_t n_ref, const itx_dev_opts &o) {
    itx_tuple T; T.start = 0; T.end = 0; T.info = ITX_CHROM_NONE; T.rec_off = rec_off;
    int32_t tid = (int32_t)x[1], pos = (int32_t)x[2];
    uint32_t mapq = (x[3] >> 8) & 0xff, lq = x[3] & 0xff, flag = x[4] >> 16, nc = x[4] & 0xffff;
    int32_t lseq = (int32_t)x[5], mpos = (int32_t)x[7], isize = (int32_t)x[8];
    bool paired = flag & 1, read1 = flag & 64;
    if (paired && !read1 && !o.treat) T.info |= ITX_F_SLOT2;
    if (flag & 4) return T;                                           /* unmapped */
    T.info |= ITX_F_MAPPED;
    if (tid < 0 || tid >= n_ref) return T;
    itx_tidinfo ti = tidtab[tid];
    if (ti.flags & ITX_TID_GLSKIP) return T;
    if (ti.flags & ITX_TID_UNKNOWN) { T.info |= ITX_F_UNKNOWN; T.start = (uint32_t)tid; return T; }
    T.info |= ITX_F_USED;
    uint32_t cend = ti.cend, start = 0, end = 0; bool minus = false, se_like = false;
    bool uniq = mapq >= o.mapQ;
    if (o.treat) se_like = true;
    else if (paired) {
        if (!(flag & 8)) {
            if (!read1) return T;
            uint32_t ai = isize < 0 ? (uint32_t)0 - (uint32_t)isize : (uint32_t)isize;
            if (ai > o.iSize || isize == 0) return T;
            if (isize > 0) { start = (uint32_t)pos; end = itx_umin(cend, start + (uint32_t)isize); }
            else { start = (uint32_t)mpos; minus = true; end = itx_umin(cend, start - (uint32_t)isize); }
        } else { if (o.discardWrongEnd) return T; se_like = true; }
    } else se_like = true;
    if (se_like) {
        start = (uint32_t)pos;
        minus = (flag & 16) != 0;
        uint32_t tmpend;
        if (nc) {
            tmpend = (uint32_t)pos;
            /* bam_calend is only observable when the end survives the extension step */
            if (o.extension == 0 || minus) {
                uint64_t cp = p + 36 + lq;
                /* only the CIGAR words that lie inside the record: an n_cigar that overstates what the record holds (a corrupt
                 * file, or the garbage a wrongly guessed span walks) must not send the loop past the record's own end */
                const uint64_t rec_end = p + 4 + (uint64_t)x[0];
                if (cp + 4ull * nc > rec_end) nc = rec_end > cp ? (uint32_t)((rec_end - cp) >> 2) : 0u;
                for (uint32_t k = 0; k < nc; k++) {
                    uint32_t cg = S.u32(cp + 4ull * k), op = cg & 0xf;
                    if (op == 0 || op == 2 || op == 3) tmpend += cg >> 4;
                }
            }
        } else tmpend = (uint32_t)(pos + lseq);
        end = itx_umin(cend, tmpend);
        if (o.extension) {
            if (!minus) end = itx_umin(start + o.extension, cend);
            else start = (end < o.extension) ? 0u : end - o.extension;
        }
    }
    T.start = start; T.end = end;
    T.info = (T.info & ~ITX_CHROM_MASK) | ITX_F_FRAG | (uniq ? ITX_F_UNIQ : 0u) | (minus ? ITX_F_MINUS : 0u) |
             (ti.chrom < 0 ? ITX_CHROM_NONE : (uint32_t)ti.chrom);
    if (WANT_XA && o.diffSubfam && ti.chrom >= 0) {
        uint64_t a0, aend; itx_aux_range(p, x, &a0, &aend);
        if (itx_aux_find(S, a0, aend, 'X', 'A')) T.info |= ITX_F_HASXA;
    }
    return T;
}

/* ------------------------------------------------------------------ interval lookup */
/* list-order key of binKeeperFind's result: level 5->0, bin high->low, rmsk row old->new.  The level is
 * the smallest l with (s >> (17+3l)) == ((e-1) >> (17+3l)), i.e. ceil(bitlength((s ^ (e-1)) >> 17) / 3). */
ITX_HD uint32_t itx_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
ITX_HD uint64_t itx_order_key(int32_t s, int32_t e, uint32_t row) {
    const uint32_t x = ((uint32_t)s ^ (uint32_t)(e - 1)) >> 17;
    uint32_t l = (32u - itx_clz32(x) + 2u) / 3u;
    if (l > 5u) l = 5u;
    const uint32_t a = l >= 5u ? 0u : (uint32_t)s >> (17u + 3u * l);      /* a shift by 32 is undefined; level 5 has one bin */
    return ((uint64_t)(5u - l) << 48) | ((uint64_t)(0xffffu - a) << 32) | (uint64_t)row;
}
ITX_HD itx_iv itx_ld_iv(const itx_dev_index &D, uint32_t i) {
#if defined(__CUDA_ARCH__)
    const int4 v = __ldg(reinterpret_cast<const int4 *>(D.iv + i));
    itx_iv e; e.start = v.x; e.end = v.y; e.pmax = v.z; e.row = (uint32_t)v.w;
    return e;
#else
    return D.iv[i];
#endif
}
/* The candidates of the clamped query [fs, fe) on a chromosome are walked DOWNWARD from the end of fe's
 * position bucket: elements above that have start >= fe; the walk skips the few in the bucket itself that
 * still start at or after fe and stops as soon as the running maximum of `end` drops to fs or below. */
struct itx_query { int32_t fs, fe; uint32_t lo, top; };      /* walk i = top-1 .. lo */
ITX_HD bool itx_query_open(const itx_dev_index &D, int32_t c, uint32_t start, uint32_t end, itx_query *q) {
#if defined(__CUDA_ARCH__)
    const int4 v = __ldg(reinterpret_cast<const int4 *>(D.cinfo + c));
    const uint32_t off = (uint32_t)v.x, bucket = (uint32_t)v.y; const int32_t size = v.z;
#else
    const uint32_t off = D.cinfo[c].off, bucket = D.cinfo[c].bucket; const int32_t size = D.cinfo[c].size;
#endif
    int32_t a = (int32_t)start, b = (int32_t)end;
    if (a < 0) a = 0;
    if (b > size) b = size;
    if (!(a < b)) return false;
    q->fs = a; q->fe = b; q->lo = off;
    q->top = D.bucket[bucket + ((uint32_t)b >> ITX_BSH) + 1];
    return true;
}
ITX_HD float itx_cov(uint32_t start, uint32_t end, int32_t es, int32_t ee) {
    int32_t s = (int32_t)start > es ? (int32_t)start : es, e = (int32_t)end < ee ? (int32_t)end : ee;
    int32_t r = e - s; if (r < 0) r = 0;
    if (r != 0 && (uint32_t)r == end - start) return 1.0f;            /* fragment inside the element: x / x == 1 exactly */
    float den = (float)(end - start);
    return den == 0.0f ? 0.0f : (float)r / den;
}
/* the overlap of [start, end) with [es, ee) as getCov's numerator takes it (generic.c:296-301) */
ITX_HD uint32_t itx_ovl(uint32_t start, uint32_t end, int32_t es, int32_t ee) {
    const int32_t s = (int32_t)start > es ? (int32_t)start : es, e = (int32_t)end < ee ? (int32_t)end : ee;
    const int32_t r = e - s;
    return r > 0 ? (uint32_t)r : 0u;
}
#define ITX_COV_FLOOR 0.0001220703125f       /* 2^-13 */
/* Coverage r / den for a caller that only asks `cov < thr` (generic.c:961): the float quotient itx_cov forms, or -- when
 * the overlap is at least 2^-12 of the fragment and thr <= 2^-13 -- just 2^-13, which lies on the same side of thr
 * (the quotient of the two rounded floats is then >= 2^-12 * (1 - 2^-22) > 2^-13 >= thr): no division. */
ITX_HD float itx_cov_thr(uint32_t r, uint32_t den, float thr) {
    if (r != 0 && r == den) return 1.0f;
    if (thr <= ITX_COV_FLOOR && den != 0 && r >= (den >> 12) + ((den & 0xfffu) ? 1u : 0u)) return ITX_COV_FLOOR;      /* r * 2^12 >= den */
    const float d = (float)den;
    return d == 0.0f ? 0.0f : (float)(int32_t)r / d;
}
/* n > 1 hits: walk the list in binKeeper order (key ascending) without materialising it -- once per list
 * position the candidates are re-walked for the smallest key above the previous one -- and apply the
 * "last ascent" rule.  Rare (nested / abutting repeats), so it is kept out of line. */
struct itx_sel_cov { long long sel; float cov; };             /* by value: nothing of the caller's goes through local memory */
ITX_HDN itx_sel_cov itx_select_multi(const itx_dev_index &D, const itx_query Q, uint32_t start, uint32_t end, int32_t n) {
    const int32_t fs = Q.fs, fe = Q.fe;
    uint64_t prev_key = 0; bool have_prev = false; float prev_cov = 0.0f, best_cov = 0.0f; long long sel = -1;
    for (int32_t k = 0; k < n; k++) {
        uint64_t bk = ~0ull; long long bi = -1; itx_iv be; be.start = be.end = 0; be.pmax = 0; be.row = 0;
        for (uint32_t i = Q.top; i-- > Q.lo;) {
            const itx_iv e = itx_ld_iv(D, i);
            if (!(e.pmax > fs)) break;
            if (e.end > fs && e.start < fe && e.start < e.end) {
                const uint64_t key = itx_order_key(e.start, e.end, e.row);
                if ((!have_prev || key > prev_key) && key < bk) { bk = key; bi = i; be = e; }
            }
        }
        if (bi < 0) break;
        const float cov = itx_cov(start, end, be.start, be.end);
        if (cov > prev_cov) { sel = bi; best_cov = cov; }
        prev_cov = cov; prev_key = bk; have_prev = true;
    }
    itx_sel_cov r; r.sel = sel; r.cov = best_cov;
    return r;
}
/* Overlap + "last ascent" selection for the fragment [start, end) on rmsk chromosome c.  Returns the
 * sorted-table index of the selected element or -1; *n_hits = length of binKeeperFind's hit list, *sel_iv = the
 * element.  Up to four hits are kept in registers (the newest with its element, the others by index) and visited in
 * list order (order key ascending); longer lists take itx_select_multi.  Exact for any n.
 * *tcov: every caller only asks `*tcov < thr` (generic.c:961), so *tcov is the selected element's coverage OR, when a
 * single hit covers at least 2^-12 of the fragment and thr <= 2^-13, just 2^-13: the same side of thr, without the
 * float division (the float quotient of the two rounded floats is then >= 2^-12 * (1 - 2^-22) > 2^-13 >= thr). */
/* the table loads of a walk: straight from global memory, or (k_scan) through a warp's shared-memory window of the table */
struct itx_iv_global {
    const itx_dev_index &D;
    ITX_HDM itx_iv operator()(uint32_t i) const { return itx_ld_iv(D, i); }
};
/* More than four hits: ITX_SEL_LONG is returned and the caller runs itx_select_multi (out of line, in a region of its own). */
#define ITX_SEL_LONG (-2ll)
template <class LdIv>
ITX_HD long long itx_select_walk(const itx_dev_index &D, const itx_query &Q, const LdIv &ld, uint32_t start, uint32_t end,
                                 float thr, int32_t *n_hits, float *tcov, itx_iv *sel_iv) {
    const int32_t fs = Q.fs, fe = Q.fe;
    /* the last four hits: the newest with its element, the others by index (a hit shifts them down: eight moves) */
    int32_t n = 0; uint32_t i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    itx_iv e0; e0.start = e0.end = 0; e0.pmax = 0; e0.row = 0;
    for (uint32_t i = Q.top; i-- > Q.lo;) {
        const itx_iv e = ld(i);
        if (!(e.pmax > fs)) break;
        if (e.end > fs && e.start < fe && e.start < e.end) { i3 = i2; i2 = i1; i1 = i0; i0 = i; e0 = e; n++; }
    }
    *n_hits = n;
    if (n == 0) return -1;
    const uint32_t den = end - start;
    if (n == 1) { *tcov = itx_cov_thr(itx_ovl(start, end, e0.start, e0.end), den, thr); *sel_iv = e0; return (long long)i0; }
    if (n > 4) return ITX_SEL_LONG;
    /* Two to four hits, visited in list order (key ascending), "last ascent" on the coverage.  All coverages share the
     * denominator, and for fragments shorter than 2^23 bases the float quotients order exactly like the overlaps
     * (both operands convert exactly, two quotients differ by at least 1 / den > 2^-23, more than one unit in the last
     * place anywhere in (0, 1]): the overlaps are compared as integers and no quotient is formed.  Longer fragments
     * take the float comparison itself. */
    const bool by_int = den != 0 && den < (1u << 23);
    const itx_iv e1 = ld(i1);
    if (n == 2) {
        const bool first0 = itx_order_key(e0.start, e0.end, e0.row) < itx_order_key(e1.start, e1.end, e1.row);
        const uint32_t r0 = itx_ovl(start, end, e0.start, e0.end), r1 = itx_ovl(start, end, e1.start, e1.end);
        bool take_b;                                        /* the second is taken only if it covers more than the first (which covers > 0) */
        if (by_int) take_b = first0 ? r1 > r0 : r0 > r1;
        else {
            const float c0 = itx_cov(start, end, e0.start, e0.end), c1 = itx_cov(start, end, e1.start, e1.end);
            take_b = first0 ? c1 > c0 : c0 > c1;
        }
        const bool pick0 = first0 != take_b;
        *tcov = itx_cov_thr(pick0 ? r0 : r1, den, thr);
        *sel_iv = pick0 ? e0 : e1;
        return (long long)(pick0 ? i0 : i1);
    }
    const itx_iv e2 = ld(i2);
    itx_iv e3 = e0; if (n > 3) e3 = ld(i3);
    const uint64_t NOKEY = ~0ull;
    uint64_t k0 = itx_order_key(e0.start, e0.end, e0.row), k1 = itx_order_key(e1.start, e1.end, e1.row);
    uint64_t k2 = itx_order_key(e2.start, e2.end, e2.row), k3 = n > 3 ? itx_order_key(e3.start, e3.end, e3.row) : NOKEY;
    const uint32_t r0 = itx_ovl(start, end, e0.start, e0.end), r1 = itx_ovl(start, end, e1.start, e1.end);
    const uint32_t r2 = itx_ovl(start, end, e2.start, e2.end), r3 = n > 3 ? itx_ovl(start, end, e3.start, e3.end) : 0u;
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
    if (!by_int) {
        c0 = itx_cov(start, end, e0.start, e0.end); c1 = itx_cov(start, end, e1.start, e1.end);
        c2 = itx_cov(start, end, e2.start, e2.end); c3 = n > 3 ? itx_cov(start, end, e3.start, e3.end) : 0.0f;
    }
    float prev = 0.0f; uint32_t prev_r = 0; int32_t sel = -1;
    for (int32_t step = 0; step < n; step++) {
        /* the unvisited hit with the smallest key */
        int32_t j = 0; uint64_t km = k0;
        if (k1 < km) { km = k1; j = 1; }
        if (k2 < km) { km = k2; j = 2; }
        if (k3 < km) { km = k3; j = 3; }
        const uint32_t rj = j == 0 ? r0 : (j == 1 ? r1 : (j == 2 ? r2 : r3));
        const float cj = j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3));
        if (by_int ? rj > prev_r : cj > prev) sel = j;
        prev = cj; prev_r = rj;
        if (j == 0) k0 = NOKEY; else if (j == 1) k1 = NOKEY; else if (j == 2) k2 = NOKEY; else k3 = NOKEY;
    }
    if (sel < 0) { *tcov = 0.0f; return -1; }
    *tcov = itx_cov_thr(sel == 0 ? r0 : (sel == 1 ? r1 : (sel == 2 ? r2 : r3)), den, thr);
    *sel_iv = sel == 0 ? e0 : (sel == 1 ? e1 : (sel == 2 ? e2 : e3));
    return (long long)(sel == 0 ? i0 : (sel == 1 ? i1 : (sel == 2 ? i2 : i3)));
}
ITX_HD long long itx_find_select(const itx_dev_index &D, int32_t c, uint32_t start, uint32_t end, float thr, int32_t *n_hits, float *tcov, itx_iv *sel_iv) {
    *n_hits = 0; *tcov = 0.0f;
    itx_query Q;
    if (!itx_query_open(D, c, start, end, &Q)) return -1;
    long long sel = itx_select_walk(D, Q, itx_iv_global{D}, start, end, thr, n_hits, tcov, sel_iv);
    if (sel == ITX_SEL_LONG) {
        const itx_sel_cov r = itx_select_multi(D, Q, start, end, *n_hits);
        *tcov = r.cov; sel = r.sel;
        if (sel >= 0) *sel_iv = itx_ld_iv(D, (uint32_t)sel);
    }
    return sel;
}
/* head of binKeeperFind's list (cpgBedGraphOverlapRepeat, generic.c:1086-1089) */
ITX_HD long long itx_find_head(const itx_dev_index &D, int32_t c, uint32_t start, uint32_t end, itx_iv *sel_iv) {
    itx_query Q;
    if (!itx_query_open(D, c, start, end, &Q)) return -1;
    const int32_t fs = Q.fs, fe = Q.fe;
    uint64_t bk = ~0ull; long long bi = -1;
    for (uint32_t i = Q.top; i-- > Q.lo;) {
        const itx_iv e = itx_ld_iv(D, i);
        if (!(e.pmax > fs)) break;
        if (e.end > fs && e.start < fe && e.start < e.end) {
            const uint64_t key = itx_order_key(e.start, e.end, e.row);
            if (key < bk) { bk = key; bi = i; *sel_iv = e; }
        }
    }
    return bi;
}
/* does [s, e) on chromosome c touch any element whose case-folded subfamily differs from `fold`? */
#ifndef ITX_XA_WALK_NOTE
#define ITX_XA_WALK_NOTE(c, s, e, fold)      /* tests/emu notes what the parsers of an alternate ask the table for */
#endif
ITX_HD bool itx_any_other_subfam(const itx_dev_index &D, int32_t c, int32_t s, int32_t e, int32_t fold) {
    ITX_XA_WALK_NOTE(c, s, e, fold);
    itx_query Q;
    if (!itx_query_open(D, c, (uint32_t)s, (uint32_t)e, &Q)) return false;
    const int32_t fs = Q.fs, fe = Q.fe;
    for (uint32_t i = Q.top; i-- > Q.lo;) {
        /* ivf: the interval with its folded subfamily id beside it -- one load per candidate instead of three dependent ones
         * (iv, meta, sinfo), and a third of the bytes: the alternates land anywhere in the table, every load is a miss */
#if defined(__CUDA_ARCH__)
        const int4 w = __ldg(reinterpret_cast<const int4 *>(D.ivf + i));
        itx_iv v; v.start = w.x; v.end = w.y; v.pmax = w.z; v.row = (uint32_t)w.w;
#else
        const itx_iv v = D.ivf[i];
#endif
        if (!(v.pmax > fs)) break;
        if (v.end > fs && v.start < fe && v.start < v.end && (int32_t)v.row != fold) return true;
    }
    return false;
}

/* ------------------------------------------------------------------ XA:Z alternates (mapped2diffSubfam) */
/* strtol(s, 0, 0) over the bytes [s, e): white space, sign, 0x / 0 prefixes, saturating; returned as (int) */
template <class Src, class Off>
ITX_HD int32_t itx_strtol_int(const Src &S, Off s, Off e) {
    while (s < e) { uint8_t c = S.u8(s); if (c == ' ' || (c >= 9 && c <= 13)) s++; else break; }
    bool neg = false;
    if (s < e) { uint8_t c = S.u8(s); if (c == '-') { neg = true; s++; } else if (c == '+') s++; }
    uint32_t base = 10;
    if (s < e && S.u8(s) == '0') {
        uint8_t c1 = s + 1 < e ? S.u8(s + 1) : 0, c2 = s + 2 < e ? S.u8(s + 2) : 0;
        bool hex2 = (c2 >= '0' && c2 <= '9') || ((c2 | 32) >= 'a' && (c2 | 32) <= 'f');
        if ((c1 == 'x' || c1 == 'X') && hex2) { base = 16; s += 2; } else base = 8;
    }
    if (base == 10) {
        /* the usual case -- at most nine decimal digits -- in 32 bits; a tenth digit goes to the general loop below, from the start */
        uint32_t a32 = 0, nd = 0; Off t = s;
        while (t < e && nd < 9u) { const uint32_t d = (uint32_t)S.u8(t) - (uint32_t)'0'; if (d > 9u) break; a32 = a32 * 10u + d; t++; nd++; }
        if (!(t < e && (uint32_t)S.u8(t) - (uint32_t)'0' <= 9u)) return neg ? (int32_t)(0u - a32) : (int32_t)a32;
    }
    uint64_t acc = 0; bool sat = false;
    while (s < e) {
        uint8_t c = S.u8(s); uint32_t d;
        if (c >= '0' && c <= '9') d = c - '0'; else if ((c | 32) >= 'a' && (c | 32) <= 'z') d = (uint32_t)((c | 32) - 'a') + 10; else break;
        if (d >= base) break;
        /* the exact overflow test needs a 64-bit division: only once the value is within a factor 36 of 2^64 */
        if (acc >= (1ull << 58) && acc > (0xffffffffffffffffull - d) / base) sat = true; else acc = acc * base + d;
        s++;
    }
    /* LONG_MIN / LONG_MAX on overflow, like glibc; the caller's (int) cast keeps the low 32 bits */
    uint64_t v;
    if (neg) v = (sat || acc > 0x8000000000000000ull) ? 0x8000000000000000ull : 0ull - acc;
    else v = (sat || acc > 0x7fffffffffffffffull) ? 0x7fffffffffffffffull : acc;
    return (int32_t)(uint32_t)v;
}
template <class Src, class Off>
ITX_HD int32_t itx_chrom_by_name(const itx_dev_index &D, const Src &S, Off s, Off e) {
    uint32_t h = 2166136261u;
    for (Off i = s; i < e; i++) { h ^= S.u8(i); h *= 16777619u; }
    uint32_t m = D.cname_nslot - 1, i = h & m;
    for (;;) {
        uint32_t v = D.cname_slot[i];
        if (!v) return -1;
        const char *nm = D.cname_pool + D.cname_off[v - 1];
        Off k = 0; bool same = true;
        for (; s + k < e; k++) if ((uint8_t)nm[k] != S.u8(s + k) || nm[k] == 0) { same = false; break; }
        if (same && nm[k] == 0) return (int32_t)(v - 1);
        i = (i + 1) & m;
    }
}
/* ---- the pieces of mapped2diffSubfam (generic.c:303-341), shared by the one-lane walk and k_scan's lane-per-alternate walk ----
 * The value of XA:Z is chopped at ';' into at most 100 pieces (chopByChar(..., 100)); piece k is what lies between the k-th and the
 * (k+1)-th ';' (the string's start / end standing in for the missing ones), an empty piece is skipped, and every other piece is
 * chopped at ',' into chr, pos, cigar, nm (the reference asserts on anything but four fields: counted in *malformed here). */
/* one piece [ps, pe), ps < pe: does this alternate (nm2 <= nm) touch an element of another folded subfamily? */
template <class Src, class Off>
ITX_HD bool itx_xa_piece(const itx_dev_index &D, const Src &S, Off ps, Off pe, int32_t nm, int32_t qlen, int32_t sel_fold, bool *malformed) {
    /* up to 4 comma separated fields (chr, pos, cigar, nm); the 4th stops at the next comma */
    Off f0s = 0, f0e = 0, f1s = 0, f1e = 0, f3s = 0, f3e = 0, q = ps; int nf = 0; bool more = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (more) {
            const Off b0 = q;
            while (q < pe && S.u8(q) != ',') q++;
            if (k == 0) { f0s = b0; f0e = q; } else if (k == 1) { f1s = b0; f1e = q; } else if (k == 3) { f3s = b0; f3e = q; }
            nf++;
            if (q >= pe) more = false; else q++;
        }
    }
    *malformed = nf != 4;
    if (nf != 4) return false;
    const int32_t nm2 = itx_strtol_int(S, f3s, f3e);
    if (nm2 > nm) return false;
    int32_t st = itx_strtol_int(S, f1s, f1e);
    if (st < 0) st = (int32_t)(0u - (uint32_t)st);
    const int32_t en = (int32_t)((uint32_t)st + (uint32_t)qlen);
    const int32_t c = itx_chrom_by_name(D, S, f0s, f0e);
    return c >= 0 && itx_any_other_subfam(D, c, st, en, sel_fold);
}
/* four bytes at a time: 0x80 in every byte of w that equals the byte replicated in c4 (exact, no carries between bytes) */
ITX_HD uint32_t itx_eq4(uint32_t w, uint32_t c4) {
    const uint32_t x = w ^ c4;
    return ~((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x)) & 0x80808080u;
}
ITX_HD uint32_t itx_popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
ITX_HD uint32_t itx_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll((unsigned long long)x);
#else
    return (uint32_t)__builtin_popcountll(x);
#endif
}
ITX_HD uint32_t itx_ctz64(uint64_t x) {          /* x != 0 */
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)x) - 1u;
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}
/* n decimal digits in the low bytes of (t0, t1, t2), as strtol(.., 0, 0) reads them when n is 1..9 and there is no leading zero (a
 * leading zero means octal or hexadecimal there): *out = the value; false: not of that shape */
ITX_HD bool itx_dec9(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t n, uint32_t *out) {
    uint32_t v = 0; bool ok = n >= 1u && n <= 9u;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t word = i < 4 ? t0 : (i < 8 ? t1 : t2);
        const uint32_t d = ((word >> (8 * (i & 3))) & 0xffu) - (uint32_t)'0';
        if ((uint32_t)i < n) { ok = ok && d <= 9u; v = v * 10u + d; }
    }
    if (n > 1u && (t0 & 0xffu) == (uint32_t)'0') ok = false;
    *out = v;
    return ok;
}
/* itx_xa_piece for a piece that lies in STAGED bytes (offsets 32 bits wide, words readable up to 47 bytes past ps): the usual
 * shape -- at most 44 bytes, "name,[+-]digits,cigar,digits" with numbers of at most nine digits without leading zeros and a name of
 * at most 31 bytes -- is taken in registers: the piece's words are loaded side by side (no load waits for the byte before it),
 * the commas found four bytes at a time, the numbers read out of registers, the name compared eight words at a time with the
 * zero-padded names of D.cname32.  Anything else is itx_xa_piece's business; the verdicts are the same by construction (and
 * checked read by read in tests/emu). */
#ifndef ITX_XA_FAST_NOTE
#define ITX_XA_FAST_NOTE(i)              /* tests/emu counts how many pieces took the register path (0) and the general one (1) */
#endif
template <class Src>
ITX_HD bool itx_xa_piece_fast(const itx_dev_index &D, const Src &S, uint32_t ps, uint32_t pe, int32_t nm, int32_t qlen, int32_t sel_fold, bool *malformed) {
    const uint32_t len = pe - ps;
    if (len <= 44u && D.cname32 != nullptr) {
        const uint32_t a = ps & ~3u, sh = (ps & 3u) * 8u;
        uint32_t r[12], w[11];
#pragma unroll
        for (int k = 0; k < 12; k++) r[k] = S.w32(a + 4u * (uint32_t)k);
#pragma unroll
        for (int k = 0; k < 11; k++) w[k] = itx_funnel_r(r[k], r[k + 1], sh);
        uint64_t cm = 0;                                           /* bit i: byte i of the piece is a comma */
#pragma unroll
        for (int k = 0; k < 11; k++) {
            if (4u * (uint32_t)k >= len) break;
            const uint32_t x = itx_eq4(w[k], 0x2c2c2c2cu) >> 7;
            cm |= (uint64_t)((x | x >> 7 | x >> 14 | x >> 21) & 0xfu) << (4 * k);
        }
        cm &= (1ull << len) - 1ull;
        if (itx_popc64(cm) >= 3u) {
            const uint32_t c1 = itx_ctz64(cm); cm &= cm - 1ull;
            const uint32_t c2 = itx_ctz64(cm); cm &= cm - 1ull;
            const uint32_t c3 = itx_ctz64(cm); cm &= cm - 1ull;
            const uint32_t c4 = cm ? itx_ctz64(cm) : len;
            /* the fourth field: the alternate's edit distance */
            uint32_t nm2 = 0, st = 0;
            const uint32_t f3 = ps + c3 + 1u, f3a = f3 & ~3u, f3s = (f3 & 3u) * 8u;
            const uint32_t u0 = S.w32(f3a), u1 = S.w32(f3a + 4u);
            bool ok = c4 - c3 - 1u <= 4u && itx_dec9(itx_funnel_r(u0, u1, f3s), 0u, 0u, c4 - c3 - 1u, &nm2);
            /* the second: the position, with its strand sign (abs() drops it) */
            const uint32_t f1 = ps + c1 + 1u, f1a = f1 & ~3u, f1s = (f1 & 3u) * 8u;
            const uint32_t v0 = S.w32(f1a), v1 = S.w32(f1a + 4u), v2 = S.w32(f1a + 8u), v3 = S.w32(f1a + 12u);
            uint32_t t0 = itx_funnel_r(v0, v1, f1s), t1 = itx_funnel_r(v1, v2, f1s), t2 = itx_funnel_r(v2, v3, f1s);
            uint32_t n1 = c2 - c1 - 1u;
            const uint32_t sg = t0 & 0xffu;
            if (n1 >= 1u && (sg == (uint32_t)'+' || sg == (uint32_t)'-')) { t0 = itx_funnel_r(t0, t1, 8u); t1 = itx_funnel_r(t1, t2, 8u); t2 >>= 8; n1--; }
            ok = ok && itx_dec9(t0, t1, t2, n1, &st);
            ok = ok && c1 >= 1u && c1 <= 31u;
            if (ok) {
                ITX_XA_FAST_NOTE(0);
                *malformed = false;
                if ((int32_t)nm2 > nm) return false;
                const int32_t en = (int32_t)(st + (uint32_t)qlen);
                /* the first: the chromosome, by name */
                uint32_t h = 2166136261u;
#pragma unroll
                for (int i = 0; i < 31; i++) { if ((uint32_t)i >= c1) break; h ^= (w[i >> 2] >> (8 * (i & 3))) & 0xffu; h *= 16777619u; }
                /* the words of the name: whole ones as they are, the last one cut to the name's bytes; the padded name in the table must
                 * equal them and be all zeros from there on (its first word past them is: names hold no zero byte) */
                const uint32_t nwd = (c1 + 3u) >> 2;               /* 1 .. 8 */
                const uint32_t m = D.cname_nslot - 1u;
                for (uint32_t i = h & m;; i = (i + 1u) & m) {
                    const uint32_t v = D.cname_slot[i];
                    if (!v) return false;                          /* no such chromosome */
                    const uint32_t *nw = D.cname32 + 8u * (v - 1u);
                    uint32_t nv[8];
#if defined(__CUDA_ARCH__)
                    const uint4 na = __ldg(reinterpret_cast<const uint4 *>(nw)), nb = __ldg(reinterpret_cast<const uint4 *>(nw) + 1);
                    nv[0] = na.x; nv[1] = na.y; nv[2] = na.z; nv[3] = na.w; nv[4] = nb.x; nv[5] = nb.y; nv[6] = nb.z; nv[7] = nb.w;
#else
                    for (int k = 0; k < 8; k++) nv[k] = nw[k];
#endif
                    bool same = true;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if ((uint32_t)k > nwd) break;
                        uint32_t want = 0u;                        /* k == nwd: the word behind the name */
                        if ((uint32_t)k < nwd) want = 4u * (uint32_t)k + 4u <= c1 ? w[k] : w[k] & ((1u << (8u * (c1 & 3u))) - 1u);
                        same = same && nv[k] == want;
                    }
                    if (same) return itx_any_other_subfam(D, (int32_t)(v - 1u), (int32_t)st, en, sel_fold);
                }
            }
        }
    }
    ITX_XA_FAST_NOTE(1);
    return itx_xa_piece(D, S, ps, pe, nm, qlen, sel_fold, malformed);
}
/* the value string that starts at zs (the byte after the type byte): *ze = its terminator (or aend), returns the number of
 * pieces the reference visits (0 for the empty string: chopByChar on "" gives none).  Aligned words once the first odd bytes are
 * done (reads may touch up to 3 bytes past aend, inside the word that holds aend - 1: stream buffers carry slack). */
template <class Src, class Off>
ITX_HD uint32_t itx_xa_count(const Src &S, Off zs, Off aend, Off *ze) {
    Off z = zs; uint32_t semis = 0; bool open = true;
    while (open && z < aend && (z & 3)) { const uint8_t c = S.u8(z); if (c == 0) open = false; else { semis += c == ';' ? 1u : 0u; z++; } }
    while (open && z + 4 <= aend) {
        const uint32_t w = S.w32(z);
        if (itx_eq4(w, 0u)) break;                               /* the terminator is in this word: the byte loop finds it */
        semis += itx_popc32(itx_eq4(w, 0x3b3b3b3bu));
        z += 4;
    }
    while (open && z < aend) { const uint8_t c = S.u8(z); if (c == 0) open = false; else { semis += c == ';' ? 1u : 0u; z++; } }
    *ze = z;
    if (z == zs) return 0;
    return semis + 1u < 100u ? semis + 1u : 100u;
}
/* itx_xa_count that also notes where the first eight ';' are: their offsets from zs, eight bits each, in sp[0..1] (*packed = false if
 * one of them does not fit eight bits) -- the lanes that take the pieces then need not scan the string again (itx_xa_piece_bounds) */
template <class Src, class Off>
ITX_HD uint32_t itx_xa_count_pack(const Src &S, Off zs, Off aend, Off *ze, uint32_t sp[2], bool *packed) {
    Off z = zs; uint32_t semis = 0; bool open = true, ok = true;
    sp[0] = sp[1] = 0;
#define ITX_XA_NOTE(pos_) do { const Off o_ = (pos_) - zs; if (semis < 8u) { if (o_ < 256u) sp[semis >> 2] |= (uint32_t)o_ << (8u * (semis & 3u)); else ok = false; } semis++; } while (0)
    while (open && z < aend && (z & 3)) { const uint8_t c = S.u8(z); if (c == 0) open = false; else { if (c == ';') ITX_XA_NOTE(z); z++; } }
    while (open && z + 4 <= aend) {
        const uint32_t w = S.w32(z);
        if (itx_eq4(w, 0u)) break;                               /* the terminator is in this word: the byte loop finds it */
        uint32_t m = itx_eq4(w, 0x3b3b3b3bu);
        while (m) {                                              /* lowest set byte first: bytes are in memory order (little endian) */
            const uint32_t low = m & (0u - m);
            const uint32_t b = low == 0x80u ? 0u : (low == 0x8000u ? 1u : (low == 0x800000u ? 2u : 3u));
            ITX_XA_NOTE(z + b);
            m &= m - 1u;
        }
        z += 4;
    }
    while (open && z < aend) { const uint8_t c = S.u8(z); if (c == 0) open = false; else { if (c == ';') ITX_XA_NOTE(z); z++; } }
#undef ITX_XA_NOTE
    *ze = z; *packed = ok;
    if (z == zs) return 0;
    return semis + 1u < 100u ? semis + 1u : 100u;
}
/* bounds of piece k out of the packed offsets (k < 8, packed); np = itx_xa_count_pack's result */
template <class Off>
ITX_HD void itx_xa_piece_bounds(Off zs, Off ze, uint32_t np, const uint32_t sp[2], uint32_t k, Off *ps, Off *pe) {
    const uint32_t prev = k ? ((k - 1u) < 4u ? sp[0] >> (8u * (k - 1u)) : sp[1] >> (8u * (k - 5u))) & 0xffu : 0u;
    const uint32_t cur = (k < 4u ? sp[0] >> (8u * k) : sp[1] >> (8u * (k - 4u))) & 0xffu;
    *ps = k ? zs + prev + 1u : zs;
    *pe = k + 1u < np ? zs + cur : ze;                          /* the last piece runs to the end of the string */
}
/* bounds of piece k (k < itx_xa_count): [*ps, *pe) lies between the k-th and the (k+1)-th ';' of [zs, ze) */
template <class Src, class Off>
ITX_HD void itx_xa_kth(const Src &S, Off zs, Off ze, uint32_t k, Off *ps, Off *pe) {
    Off z = zs; uint32_t seen = 0;
    while (seen < k && z < ze && (z & 3)) { if (S.u8(z) == ';') seen++; z++; }
    while (seen < k && z + 4 <= ze) {
        const uint32_t n = itx_popc32(itx_eq4(S.w32(z), 0x3b3b3b3bu));
        if (seen + n >= k) break;                                /* the k-th ';' is in this word */
        seen += n; z += 4;
    }
    while (seen < k && z < ze) { if (S.u8(z) == ';') seen++; z++; }
    *ps = z;
    while (z < ze && (z & 3) && S.u8(z) != ';') z++;
    if (z < ze && !(z & 3)) {
        while (z + 4 <= ze && !itx_eq4(S.w32(z), 0x3b3b3b3bu)) z += 4;
        while (z < ze && S.u8(z) != ';') z++;
    }
    *pe = z;
}
/* the one-lane walk over the alternates: xa = offset of the type byte of XA (0: no such tag), nm = bam_aux2i of NM.
 * *malformed counts alternates without 4 comma separated fields met before the verdict. */
/* S by value: an out-of-line function handed a reference would read the source's fields back from local memory at every byte */
template <class Src>
ITX_HDN bool itx_xa_walk(const itx_dev_index &D, const Src S, uint64_t xa, uint64_t aend, int32_t nm, int32_t sel_fold, int32_t qlen, uint32_t *malformed) {
    if (!xa || xa >= aend) return false;
    const uint8_t ty = S.u8(xa);
    if (ty != 'Z' && ty != 'H') return false;
    uint64_t s = xa + 1, zend = s;
    while (zend < aend && S.u8(zend) != 0) zend++;
    if (s == zend) return false;                                  /* chopByChar on "" gives no pieces */
    int pieces = 0;
    while (pieces < 100) {
        uint64_t pe = s;
        while (pe < zend && S.u8(pe) != ';') pe++;
        pieces++;
        if (pe > s) {
            bool mal;
            if (itx_xa_piece(D, S, s, pe, nm, qlen, sel_fold, &mal)) return true;
            if (mal && malformed) (*malformed)++;
        }
        if (pe >= zend) break;
        s = pe + 1;
    }
    return false;
}
/* record's aux area [a0, aend) carries XA; sel_fold = folded subfamily of the selected element; qlen = end - start */
template <class Src>
ITX_HD bool itx_mapped_to_diff_subfam_aux(const itx_dev_index &D, const Src &S, uint64_t a0, uint64_t aend,
                                          int32_t sel_fold, int32_t qlen, uint32_t *malformed) {
    const uint64_t xa = itx_aux_find(S, a0, aend, 'X', 'A');
    if (!xa || xa >= aend) return false;
    const int32_t nm = itx_aux2i(S, itx_aux_find(S, a0, aend, 'N', 'M'), aend);
    return itx_xa_walk(D, S, xa, aend, nm, sel_fold, qlen, malformed);
}

template <class Src>
ITX_HD bool itx_mapped_to_diff_subfam(const itx_dev_index &D, const Src &S, uint64_t p, const uint32_t x[9],
                                      int32_t sel_fold, int32_t qlen, uint32_t *malformed) {
    uint64_t a0, aend; itx_aux_range(p, x, &a0, &aend);
    return itx_mapped_to_diff_subfam_aux(D, S, a0, aend, sel_fold, qlen, malformed);
}

/* ------------------------------------------------------------------ consensus coverage range */
/* closed form of the loop at generic.c:990-1006: the j range [*j0, *j1) touched by a fragment
 * (start, qlen) on element (es, ee, cs, ce) of a subfamily with consensus length L; false = nothing */
ITX_HD bool itx_cov_range(uint32_t start, uint32_t qlen, int32_t es, int32_t ee, uint32_t cs, uint32_t ce, uint32_t L,
                          uint32_t *j0, uint32_t *j1) {
    uint32_t rstart = start - (uint32_t)es;
    uint32_t rend = rstart + qlen;
    rend = rend < (uint32_t)ee ? rend : (uint32_t)ee;
    if (!(rstart < rend)) return false;
    uint32_t a = rstart + cs, lim = ce < L ? ce : L;
    if (!(a < lim)) return false;
    uint32_t n = rend - rstart, room = lim - a;
    *j0 = a; *j1 = a + (n < room ? n : room);
    return true;
}
/* ------------------------------------------------------------------ -R duplicate removal (generic.c:907-919) */
/* The reference keeps a hash of "chr:start:end:strand" strings and drops a read whose key is already there.  The
 * key is only re-formatted for unique reads (MAPQ >= -Q); any other read re-uses the key left by the last unique
 * read before it -- which is in the hash by then -- so it is dropped, except while no unique read has been seen
 * yet: then the key is still the empty string, which the first such read adds.  In terms of the file-order
 * ordinal of a fragment:
 *   unique read      kept <=> no unique read with the same key has a smaller ordinal
 *   any other read   kept <=> it is the first non-unique read of the run and precedes every unique read
 * State persists across the files of a run.  The key is (chromosome identity, strand, start, end). */
ITX_HD void itx_dup_key(int32_t csid, bool minus, uint32_t start, uint32_t end, unsigned long long *lo, unsigned long long *hi) {
    *lo = (unsigned long long)start | ((unsigned long long)end << 32);
    *hi = (((unsigned long long)(uint32_t)(csid + 1)) << 1) | (minus ? 1ull : 0ull);
}
ITX_HD bool itx_dup_nonunique_kept(unsigned long long ord, unsigned long long min_unique, unsigned long long min_nonunique) {
    return ord == min_nonunique && ord < min_unique;
}
ITX_HD unsigned long long itx_dup_hash(unsigned long long lo, unsigned long long hi) {
    unsigned long long x = lo ^ (hi * 0x9E3779B97F4A7C15ull);
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    return x;
}
/* ------------------------------------------------------------------ bedGraph text (cpgBedGraphOverlapRepeat, generic.c:1069-1076) */
ITX_HD bool itx_is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
/* strtod over [s, e) for the plain decimal forms [sign] digits [. digits] [e|E [sign] digits]: at most 19 significant
 * digits gathered exactly in 64 bits; when the integer fits 53 bits and the power of ten is at most 22, ONE IEEE
 * multiplication or division of two exact doubles gives the correctly rounded value -- the value glibc's strtod
 * returns.  Anything else (longer mantissas, huge exponents, inf / nan / hex floats) sets *exact = false and the
 * caller hands the line to the host's strtod.  Trailing characters stop the parse like strtod(s, NULL). */
ITX_HD double itx_pow10_exact(uint32_t k) {       /* 10^k, k <= 22: all exactly representable */
    double r = 1.0;
    const double t[5] = {1e1, 1e2, 1e4, 1e8, 1e16};
#pragma unroll
    for (int b = 0; b < 5; b++) if (k & (1u << b)) r *= t[b];          /* products of exact powers below 2^53 * 2^k stay exact up to 1e22 */
    return r;
}
template <class Src>
ITX_HD double itx_strtod_fast(const Src &S, uint64_t s, uint64_t e, bool *exact) {
    *exact = true;
    while (s < e && itx_is_space(S.u8(s))) s++;
    bool neg = false;
    if (s < e) { const uint8_t c = S.u8(s); if (c == '-') { neg = true; s++; } else if (c == '+') s++; }
    uint64_t m = 0; int32_t e10 = 0; uint32_t nd = 0; bool any = false, dropped = false;
    while (s < e) { const uint8_t c = S.u8(s); if (c < '0' || c > '9') break; any = true; if (nd < 19) { m = m * 10 + (c - '0'); if (m) nd++; } else { e10++; if (c != '0') dropped = true; } s++; }
    if (s < e && S.u8(s) == '.') {
        s++;
        while (s < e) { const uint8_t c = S.u8(s); if (c < '0' || c > '9') break; any = true; if (nd < 19) { m = m * 10 + (c - '0'); if (m) nd++; e10--; } else if (c != '0') dropped = true; s++; }
    }
    if (!any) { *exact = false; return 0.0; }                       /* inf, nan, garbage: the host decides */
    if (s < e && (S.u8(s) == 'e' || S.u8(s) == 'E')) {
        uint64_t q = s + 1; bool en = false;
        if (q < e && (S.u8(q) == '-' || S.u8(q) == '+')) { en = S.u8(q) == '-'; q++; }
        if (q < e && S.u8(q) >= '0' && S.u8(q) <= '9') {
            int32_t x = 0;
            while (q < e && S.u8(q) >= '0' && S.u8(q) <= '9') { if (x < 100000) x = x * 10 + (S.u8(q) - '0'); q++; }
            e10 += en ? -x : x;
        }
    } else if (s < e && (S.u8(s) == 'x' || S.u8(s) == 'X') && m == 0) { *exact = false; return 0.0; }      /* 0x...: a hex float */
    if (m == 0) return neg ? -0.0 : 0.0;
    if (dropped || m >= (1ull << 53) || e10 > 22 || e10 < -22) { *exact = false; return 0.0; }
    double v = (double)m;
    v = e10 >= 0 ? v * itx_pow10_exact((uint32_t)e10) : v / itx_pow10_exact((uint32_t)(-e10));
    return neg ? -v : v;
}
#endif
