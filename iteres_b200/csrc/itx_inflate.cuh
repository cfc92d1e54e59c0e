/* itx_inflate.cuh -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for one GPU thread per block.
 *
 * BGZF (cussamtools/bgzf.c:367-397 inflate_block: inflateInit2(-15) + inflate(Z_FINISH), no CRC check)
 * stores every <= 64 KiB of the BAM stream as an independent raw-deflate stream, so the blocks of a file
 * can be inflated by tens of thousands of threads at once: the 32 lanes of a warp decode 32 different
 * blocks in lock step, one symbol per lane per round.
 *
 * Decoding is table driven: a 256-entry look-up table for the literal/length code (first 8 bits) and a
 * 64-entry one for the distance code (first 6 bits) answer ~96 % / ~99 % of the symbols of a BAM stream
 * with one load; longer codes finish with the canonical bit-serial walk from length 9 / 7 (per-length code
 * counts in registers, symbols in a per-thread array).  The two tables are 640 bytes per thread and live
 * in SHARED memory, interleaved across the lanes of the warp so that every lane owns a bank; the symbol
 * arrays of the long codes (rarely read) live in global memory.  While a dynamic header is read the same
 * shared cells hold the code-length code's table and the literal code lengths, so a header costs no
 * global-memory round trips either.
 *
 * `Tab` is the per-thread store: lut(j) / lut_set(j, v) address the thread's ITX_LUT_CELLS shared 16-bit
 * cells, tab(j) / tab.set(j, v) its ITX_T_CELLS global ones.  The test-only host build uses plain arrays.
 * Like itx_logic.cuh this file is __host__ __device__ so the non-GPU suite can check it against zlib.
 */
#ifndef ITX_INFLATE_CUH
#define ITX_INFLATE_CUH
#include "itx_logic.cuh"

#define ITX_INF_OK 0
#define ITX_INF_EDATA 1        /* invalid deflate data */
#define ITX_INF_ESIZE 2        /* output is not the ISIZE the BGZF footer promised */

#ifndef ITX_LB
#define ITX_LB 8u              /* index bits of the literal/length table (>= 7: the cells hold the code-length code's 7-bit table while a header is read) */
#endif
#define ITX_DB 6u              /* index bits of the distance table */
/* shared cells of one thread (16 bit): entry = symbol << 4 | code length, 0 = "longer than the index" */
#define ITX_LUT_L 0u           /* [256] literal/length table; while a header is read: [0,128) the code-length code's table */
#define ITX_LUT_D (1u << ITX_LB)   /* [64]  distance table; while a header is read: the 256 literal code lengths, four per cell */
#define ITX_LUT_CELLS (ITX_LUT_D + 64u)
/* global cells of one thread (16 bit) */
#define ITX_T_LSYM 0u          /* [288] literal/length symbols in canonical order */
#define ITX_T_DSYM 288u        /* [30]  distance symbols in canonical order */
#define ITX_T_CELLS 320u

#define ITX_ST_HEADER 0u
#define ITX_ST_SYMBOL 1u
#define ITX_ST_DONE 2u
#define ITX_ST_ERROR 3u
#define ITX_ST_COPY 4u         /* a match is being copied, ITX_COPY_STEP bytes per round (in-line mode only) */
#define ITX_ST_OVERFLOW 5u     /* deferred mode: the block has more matches than its list holds */
#define ITX_COPY_STEP 8u
#ifndef ITX_BURST
#define ITX_BURST 3u           /* literals decoded ahead of the general symbol of a round */
#endif
#define ITX_M_NONE 0xffffffffu /* match count of a block that overflowed or failed */

ITX_HD uint32_t itx_brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
#endif
}

template <class Tab>
struct itx_inflater {
    /* input: aligned 32-bit words, three of them always in flight ahead of the bit buffer */
    const uint32_t *inw, *in_lim;          /* next word to load; the word after the last one of the stream */
    uint32_t next0, next1, next2;
    uint64_t bitbuf; uint32_t bitcnt;
    uint8_t *out; uint32_t out_cap, out_pos;
    Tab tab;
    uint32_t err;
    uint32_t state, last, expect;
    uint32_t pend_len, pend_dist;          /* ITX_ST_COPY */
    /* deferred mode (m_cap != 0): literals go to their final place, matches are only LISTED -- entry k is
     * (output position | length << 16, distance) -- and a second pass copies them while the block's
     * history is cache resident; a decoder that copies in line waits on a DRAM read of its own history every round */
    uint32_t *m_pl; uint16_t *m_d; uint32_t n_match, m_cap;
    uint32_t lc[8], dc[8];                 /* per-length code counts, two 16-bit counts per register */
    uint32_t lfirst, lindex, dfirst, dindex;   /* canonical walk state after ITX_LB / ITX_DB bits */

    ITX_HDM static uint32_t ldw(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
        return __ldg(p);
#else
        return *p;
#endif
    }
    /* `in` may have any alignment; reads run up to 28 bytes past in + in_len (buffers carry slack) and never before
     * the aligned word holding in[0] */
    ITX_HDM void set_input(const uint8_t *in, uint32_t in_len) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(in);
        inw = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        in_lim = reinterpret_cast<const uint32_t *>((a + in_len + 3) & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        bitbuf = (uint64_t)(ldw(inw) >> sh); bitcnt = 32 - sh;
        next0 = ldw(inw + 1); next1 = ldw(inw + 2); next2 = ldw(inw + 3); inw += 4;
    }
    /* at least 32 valid bits afterwards */
    ITX_HDM void refill() {
        if (bitcnt <= 32) {
            bitbuf |= (uint64_t)next0 << bitcnt; bitcnt += 32;
            next0 = next1; next1 = next2; next2 = ldw(inw); inw++;
        }
    }
    ITX_HDM bool overrun() const { return inw > in_lim + 5; }      /* more words loaded than the stream has (+ the three in flight and the bit buffer) */
    ITX_HDM void consume(uint32_t n) { bitbuf >>= n; bitcnt -= n; }
    ITX_HDM uint32_t take(uint32_t n) { const uint32_t v = (uint32_t)bitbuf & ((1u << n) - 1u); consume(n); return v; }   /* no refill: the caller knows the bits are there */
    ITX_HDM uint32_t bits(uint32_t n) { refill(); return take(n); }                                                    /* n <= 16 */

    /* ---- code lengths while a header is read: literals (0..255) in the distance cells, four per cell; the
     * length codes (256..287) and the distance codes in registers, eight per register */
    uint32_t hi_len[4], d_len[4];
    ITX_HDM uint32_t nib_get(const uint32_t r[4], uint32_t i) const {
        const uint32_t w = (i >> 3) == 0 ? r[0] : ((i >> 3) == 1 ? r[1] : ((i >> 3) == 2 ? r[2] : r[3]));
        return (w >> ((i & 7) * 4)) & 15u;
    }
    ITX_HDM void nib_set(uint32_t r[4], uint32_t i, uint32_t v) {
        const uint32_t m = v << ((i & 7) * 4);
        if ((i >> 3) == 0) r[0] |= m; else if ((i >> 3) == 1) r[1] |= m; else if ((i >> 3) == 2) r[2] |= m; else r[3] |= m;
    }
    ITX_HDM uint32_t lit_len(uint32_t s) const { return ((uint32_t)tab.lut(ITX_LUT_D + (s >> 2)) >> ((s & 3) * 4)) & 15u; }
    /* length of literal/length symbol s (0..287) */
    ITX_HDM uint32_t ll_len(uint32_t s) const { return s < 256 ? lit_len(s) : nib_get(hi_len, s - 256); }

    /* ---- canonical code from per-symbol lengths: counts -> completeness, table fill, symbols of the long codes.
     * LenAt(s) is the length of symbol s.  Returns `left` like zlib/puff: 0 complete, <0 over-subscribed, >0 incomplete. */
    template <class LenAt>
    ITX_HDM int32_t build(uint32_t n, LenAt len_at, uint32_t lut_base, uint32_t lut_bits, uint32_t sym_base, uint32_t cpk[8],
                          uint32_t *first_out, uint32_t *index_out, uint32_t *n_short) {
        uint32_t cn[16];
#pragma unroll
        for (uint32_t l = 0; l < 16; l++) cn[l] = 0;
        for (uint32_t s = 0; s < n; s++) {
            const uint32_t l = len_at(s);
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) if (k == l) cn[k]++;
        }
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) cpk[k] = cn[2 * k] | (cn[2 * k + 1] << 16);
        *n_short = cn[0] + cn[1];
        for (uint32_t j = 0; j < (1u << lut_bits); j++) tab.lut_set(lut_base + j, 0);
        if (cn[0] == n) { *first_out = 0; *index_out = 0; return 0; }     /* no codes at all: complete, every decode fails */
        int32_t left = 1;
#pragma unroll
        for (uint32_t l = 1; l <= 15; l++) { left <<= 1; left -= (int32_t)cn[l]; if (left < 0) return left; }
        /* running next code / next symbol-array slot of each length */
        uint32_t nx[16], of[16];
        nx[0] = 0; of[0] = 0; nx[1] = 0; of[1] = 0;
#pragma unroll
        for (uint32_t l = 1; l < 15; l++) { nx[l + 1] = (nx[l] + cn[l]) << 1; of[l + 1] = of[l] + cn[l]; }
        /* the bit-serial walk resumes after lut_bits bits: first = first code of the next length, index = its slot */
        {
            uint32_t f = 0, ix = 0;
#pragma unroll
            for (uint32_t l = 1; l <= 15; l++) if (l <= lut_bits) { f += cn[l]; f <<= 1; ix += cn[l]; }
            *first_out = f; *index_out = ix;
        }
        for (uint32_t s = 0; s < n; s++) {
            const uint32_t l = len_at(s);
            if (!l) continue;
            uint32_t code = 0, o = 0;
#pragma unroll
            for (uint32_t k = 1; k < 16; k++) if (k == l) { code = nx[k]; nx[k] = code + 1; o = of[k]; of[k] = o + 1; }
            tab.set(sym_base + o, (uint16_t)s);
            if (l <= lut_bits) {
                const uint16_t e = (uint16_t)((s << 4) | l);
                for (uint32_t j = itx_brev32(code) >> (32 - l); j < (1u << lut_bits); j += 1u << l) tab.lut_set(lut_base + j, e);
            }
        }
        return left;
    }
    /* codes longer than the table index: canonical bit-serial walk from length `from` + 1.  Returns the slot of
     * the symbol in the (global-memory) symbol array, or -1. */
    ITX_HDM int32_t walk_long(const uint32_t c[8], uint32_t from, uint32_t first0, uint32_t index0) {
        uint32_t buf = (uint32_t)bitbuf;
        int32_t code = (int32_t)((itx_brev32(buf) >> (32 - from)) << 1), first = (int32_t)first0, index = (int32_t)index0;
        buf >>= from;
#pragma unroll
        for (uint32_t len = 1; len <= 15; len++) {
            if (len > from) {
                code |= (int32_t)(buf & 1u); buf >>= 1;
                const int32_t count = (int32_t)((len & 1) ? (c[len >> 1] >> 16) : (c[len >> 1] & 0xffffu));
                if (code - count < first) { consume(len); return index + (code - first); }
                index += count; first += count; first <<= 1; code <<= 1;
            }
        }
        return -1;
    }

    ITX_HDM void put(uint8_t b) { if (out_pos < out_cap) out[out_pos] = b; out_pos++; }
    /* length / distance base values and extra bits (RFC 1951 3.2.5), computed instead of tabulated */
    static ITX_HDM uint32_t lext(uint32_t s) { return (s < 8 || s == 28) ? 0u : (s - 4) >> 2; }
    static ITX_HDM uint32_t lbase(uint32_t s) { return s < 8 ? 3 + s : (s == 28 ? 258u : 3 + ((4 + (s & 3)) << ((s - 4) >> 2))); }
    static ITX_HDM uint32_t dext(uint32_t d) { return d < 4 ? 0u : (d - 2) >> 1; }
    static ITX_HDM uint32_t dbase(uint32_t d) { return d < 4 ? 1 + d : 1 + ((2 + (d & 1)) << ((d - 2) >> 1)); }

    /* up to ITX_COPY_STEP bytes of the pending match (the caller checked that it fits the output) */
    ITX_HDM void copy_step() {
        const uint32_t n = pend_len < ITX_COPY_STEP ? pend_len : ITX_COPY_STEP;
        uint8_t *o = out + out_pos; const uint8_t *f = o - pend_dist;
        if (pend_dist >= ITX_COPY_STEP) {
            /* the loads are independent of each other and of the stores: one memory latency per step */
            uint8_t t[ITX_COPY_STEP];
#pragma unroll
            for (uint32_t k = 0; k < ITX_COPY_STEP; k++) t[k] = k < n ? f[k] : (uint8_t)0;
#pragma unroll
            for (uint32_t k = 0; k < ITX_COPY_STEP; k++) if (k < n) o[k] = t[k];
        } else {
            for (uint32_t k = 0; k < n; k++) o[k] = f[k];                /* overlapping: bytes replicate, as LZ77 requires */
        }
        out_pos += n; pend_len -= n;
        state = pend_len ? ITX_ST_COPY : ITX_ST_SYMBOL;
    }
    /* One round of the symbol state: up to ITX_BURST short-code literals, then one general symbol down to its match.  Written
     * without early exits, as nested phases, so that the lanes of a warp (32 different blocks, each somewhere else in this code)
     * re-converge after every phase instead of running the tail once per path.  A code longer than the table index is finished by
     * the canonical walk and its symbol fetched from the (global-memory) symbol array on the spot.  false = invalid data. */
    ITX_HDM bool symbol_round() {
        bool ok = !(out_pos > out_cap || overrun());
        uint32_t s = 0x7fffffffu;                                          /* no symbol */
        if (ok) {
            refill();
            uint32_t e = tab.lut(ITX_LUT_L + ((uint32_t)bitbuf & ((1u << ITX_LB) - 1u)));
            /* most symbols of a BAM stream are literals with short codes: up to ITX_BURST of them go out right here,
             * before the general symbol of the round (3 x 8 of the >= 32 bits at most, so every index is made of real bits) */
#pragma unroll
            for (uint32_t b = 0; b < ITX_BURST; b++) {
                if ((e & 15u) && e < (256u << 4)) {
                    consume(e & 15u); put((uint8_t)(e >> 4));
                    e = tab.lut(ITX_LUT_L + ((uint32_t)bitbuf & ((1u << ITX_LB) - 1u)));
                }
            }
            refill();                                                      /* only adds high bits: e stays valid */
            if (e & 15u) { consume(e & 15u); s = e >> 4; }
            else {
                const int32_t slot = walk_long(lc, ITX_LB, lfirst, lindex);
                if (slot < 0) ok = false; else s = tab(ITX_T_LSYM + (uint32_t)slot);
            }
        }
        if (s < 256u) put((uint8_t)s);
        else if (s == 256u) state = last ? ITX_ST_DONE : ITX_ST_HEADER;
        else if (s != 0x7fffffffu) {
            s -= 257u;
            if (s >= 29u) ok = false;
            else {
                const uint32_t len = lbase(s) + take(lext(s));             /* <= 15 + 5 of the >= 32 bits are gone */
                refill();
                const uint32_t e = tab.lut(ITX_LUT_D + ((uint32_t)bitbuf & ((1u << ITX_DB) - 1u)));
                uint32_t d = 0xffffffffu;
                if (e & 15u) { consume(e & 15u); d = e >> 4; }
                else {
                    const int32_t slot = walk_long(dc, ITX_DB, dfirst, dindex);
                    if (slot >= 0) d = tab(ITX_T_DSYM + (uint32_t)slot);
                }
                if (d >= 30u) ok = false;
                else {
                    const uint32_t dist = dbase(d) + take(dext(d));        /* <= 15 + 13 of the >= 32 bits */
                    if (dist > out_pos) ok = false;
                    else if (out_pos + len > out_cap) { out_pos += len; ok = false; }
                    else if (m_cap) {
                        if (n_match >= m_cap) state = ITX_ST_OVERFLOW;
                        else { m_pl[n_match] = out_pos | (len << 16); m_d[n_match] = (uint16_t)dist; n_match++; out_pos += len; }
                    } else { pend_len = len; pend_dist = dist; copy_step(); }
                }
            }
        }
        return ok;
    }
    ITX_HDM bool stored() {
        consume(bitcnt & 7);                                          /* to the next byte boundary */
        const uint32_t len = bits(16), nlen = bits(16);
        if ((len ^ 0xffffu) != nlen) return false;
        for (uint32_t k = 0; k < len; k++) { put((uint8_t)bits(8)); if (overrun()) return false; }
        return true;
    }
    /* tables from the code lengths gathered by the header: literal/length first (it reads the literal lengths out
     * of the distance cells), then the distance table over those cells */
    ITX_HDM bool make_tables(uint32_t nlen, uint32_t ndist, bool check) {
        itx_inflater *self = this;
        uint32_t nshort;
        int32_t e = build(nlen, [self](uint32_t s) { return self->ll_len(s); }, ITX_LUT_L, ITX_LB, ITX_T_LSYM, lc, &lfirst, &lindex, &nshort);
        if (check && e != 0 && (e < 0 || nlen != nshort)) return false;
        e = build(ndist, [self](uint32_t s) { return self->nib_get(self->d_len, s); }, ITX_LUT_D, ITX_DB, ITX_T_DSYM, dc, &dfirst, &dindex, &nshort);
        if (check && e != 0 && (e < 0 || ndist != nshort)) return false;
        return true;
    }
    ITX_HDM void clear_lens() {
#pragma unroll
        for (int k = 0; k < 4; k++) { hi_len[k] = 0; d_len[k] = 0; }
    }
    ITX_HDM bool fixed() {
        clear_lens();
        for (uint32_t c = 0; c < 64; c++) tab.lut_set(ITX_LUT_D + c, (uint16_t)(c < 36 ? 0x8888u : 0x9999u));     /* 0..143: 8, 144..255: 9 */
        for (uint32_t s = 256; s < 288; s++) nib_set(hi_len, s - 256, s < 280 ? 7u : 8u);
        for (uint32_t s = 0; s < 30; s++) nib_set(d_len, s, 5u);
        return make_tables(288, 30, false);                   /* the fixed distance code is incomplete by definition */
    }
    ITX_HDM bool dynamic() {
        /* order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15, five bits each */
        const uint64_t O0 = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
        const uint64_t O1 = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
        const uint32_t nlen = bits(5) + 257, ndist = bits(5) + 1, ncode = bits(4) + 4;
        if (nlen > 286 || ndist > 30) return false;
        /* the 19 lengths of the code-length code, three bits each, in one 64-bit word */
        uint64_t cl = 0;
        for (uint32_t i = 0; i < ncode; i++) cl |= (uint64_t)bits(3) << (3u * (uint32_t)((i < 12 ? O0 >> (5 * i) : O1 >> (5 * (i - 12))) & 31));
        {
            uint32_t cpk[8], f, ix, nshort;
            if (build(19, [cl](uint32_t s) { return (uint32_t)(cl >> (3u * s)) & 7u; }, ITX_LUT_L, 7, ITX_T_DSYM, cpk, &f, &ix, &nshort) != 0) return false;
        }
        clear_lens();
        uint32_t i = 0, prev = 0, acc = 0;               /* acc gathers four literal lengths before they go to their cell */
        while (i < nlen + ndist) {
            refill();
            const uint32_t e = tab.lut(ITX_LUT_L + ((uint32_t)bitbuf & 127u));
            if (!(e & 15u)) return false;                /* the code-length code has no code longer than 7 bits */
            consume(e & 15u);
            const uint32_t s = e >> 4;
            uint32_t v = 0, rep = 1;
            if (s < 16) v = s;
            else if (s == 16) { if (i == 0) return false; v = prev; rep = 3 + take(2); }
            else if (s == 17) rep = 3 + take(3);
            else rep = 11 + take(7);
            if (i + rep > nlen + ndist) return false;
            for (; rep; rep--, i++) {
                if (i < 256) {
                    acc |= v << ((i & 3) * 4);
                    if ((i & 3) == 3) { tab.lut_set(ITX_LUT_D + (i >> 2), (uint16_t)acc); acc = 0; }
                } else if (i < nlen) nib_set(hi_len, i - 256, v);
                else nib_set(d_len, i - nlen, v);
            }
            prev = v;
            if (overrun()) return false;
        }
        /* nlen >= 257, so the 256 literal lengths are all in their cells */
        if (nib_get(hi_len, 0) == 0) return false;                  /* no end-of-block code */
        return make_tables(nlen, ndist, true);
    }
    /* The decoder is a small state machine so that the 32 lanes of a warp (32 different BGZF blocks) can be
     * stepped together: every call of advance() does ONE unit of work -- a block header (with its table build),
     * one symbol, or one step of a long match copy -- and the kernel re-converges the warp between calls. */
    ITX_HDM void begin(const uint8_t *in, uint32_t in_len, uint32_t expect_) {
        set_input(in, in_len);
        out_pos = 0; state = ITX_ST_HEADER; last = 0; expect = expect_; err = ITX_INF_OK; pend_len = pend_dist = 0; n_match = 0;
        lfirst = lindex = dfirst = dindex = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { lc[k] = 0; dc[k] = 0; }
        clear_lens();
    }
    ITX_HDM bool running() const { return state != ITX_ST_DONE && state != ITX_ST_ERROR && state != ITX_ST_OVERFLOW; }
    ITX_HDM void advance() {
        if (state == ITX_ST_SYMBOL) {
            if (!symbol_round()) { state = ITX_ST_ERROR; err = ITX_INF_EDATA; }
        } else if (state == ITX_ST_COPY) {
            copy_step();
        } else if (state == ITX_ST_HEADER) {
            last = bits(1);
            const uint32_t type = take(2);
            bool ok;
            if (type == 0) { ok = stored(); state = last ? ITX_ST_DONE : ITX_ST_HEADER; }
            else { ok = type == 1 ? fixed() : (type == 2 ? dynamic() : false); state = ITX_ST_SYMBOL; }
            if (!ok || overrun()) { state = ITX_ST_ERROR; err = ITX_INF_EDATA; }
        }
        if (state == ITX_ST_DONE && out_pos != expect) { state = ITX_ST_ERROR; err = ITX_INF_ESIZE; }
    }
    /* one whole raw-deflate stream; returns ITX_INF_* */
    ITX_HDM uint32_t run(const uint8_t *in, uint32_t in_len, uint32_t expect_) {
        begin(in, in_len, expect_);
        while (running()) advance();
        return err;
    }
};
/* ------------------------------------------------------------------ second pass of the deferred mode */
/* copy one match: o[0, len) = o[-d, len - d), bytes replicating when d < len as LZ77 requires.  Eight independent
 * loads per step (one memory latency per eight bytes); periods shorter than eight are replicated out of registers. */
template <bool VIA_L2>
ITX_HD uint8_t itx_lz_ld(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    if (VIA_L2) return __ldcg(p);        /* global memory: the bytes may have been stored by another lane a moment ago -- read them where stores land */
#endif
    return *p;
}
template <bool VIA_L2>
ITX_HD void itx_lz_copy(uint8_t *o, uint32_t len, uint32_t d) {
    const uint8_t *f = o - d;
    if (d >= 8u || d >= len) {
        for (uint32_t k = 0; k < len; k += 8u) {
            const uint32_t n = len - k < 8u ? len - k : 8u;
            uint8_t t[8];
#pragma unroll
            for (uint32_t j = 0; j < 8u; j++) t[j] = j < n ? itx_lz_ld<VIA_L2>(f + k + j) : (uint8_t)0;
#pragma unroll
            for (uint32_t j = 0; j < 8u; j++) if (j < n) o[k + j] = t[j];
        }
    } else {
        uint8_t t[8];
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++) t[j] = j < d ? itx_lz_ld<VIA_L2>(f + j) : (uint8_t)0;
        uint32_t idx = 0;
        for (uint32_t k = 0; k < len; k++) {
            uint8_t b = t[0];
#pragma unroll
            for (uint32_t j = 1; j < 8u; j++) if (idx == j) b = t[j];
            o[k] = b;
            idx = idx + 1u == d ? 0u : idx + 1u;
        }
    }
}
/* A batch of consecutive list entries is resolved by the lanes of a warp together.  Everything before the output
 * position of the first unfinished entry is final, so an entry may go as soon as its source ends at or before that
 * position; the first unfinished entry itself may always go. */
ITX_HD bool itx_lz_ready(uint32_t pos, uint32_t len, uint32_t dist, bool is_first_unfinished, uint32_t first_unfinished_pos) {
    return is_first_unfinished || pos - dist + len <= first_unfinished_pos;
}
/* ---- the second pass INSIDE the decoding warp: windows of W bytes.  src[i - w0] = where byte i of the window comes from (itself for
 * a literal); the chains that stay inside the window are shortened by pointer jumping, and a source before the window is final
 * because the windows of a block are resolved in order.  Only the src cells live in shared memory (the warp's look-up tables, dead by
 * then): literals and history are read where they lie.  The per-lane pieces are here so that the host test can step them lane by
 * lane; the device driver is itx_lzw_resolve (itx_kernels.cuh). */
ITX_HD void itx_lzw_init(uint16_t *src, uint32_t w0, uint32_t cnt, uint32_t lane) {
    /* up to the next multiple of 256 cells (the windows are multiples of 256): the passes below run without bound tests, and a cell
     * past the window's end is its own source like a literal (positions past 65535 wrap: such cells are never read as a source) */
    const uint32_t cnt_r = (cnt + 255u) & ~255u;
    for (uint32_t i = lane; i < cnt_r; i += 32u) src[i] = (uint16_t)(w0 + i);
}
/* the part of match (pos, len, dist) that lies in [w0, w1).  A match that overlaps itself (dist < len: a run) points every byte
 * straight at the period before the match, so the chain through its own bytes never has to be followed */
ITX_HD void itx_lzw_scatter(uint16_t *src, uint32_t w0, uint32_t w1, uint32_t pos, uint32_t len, uint32_t dist) {
    const uint32_t a = pos > w0 ? pos : w0, b = pos + len < w1 ? pos + len : w1;
    if (dist >= len) {
        for (uint32_t j = a; j < b; j++) src[j - w0] = (uint16_t)(j - dist);
    } else {
        uint32_t r = (a - pos) % dist;
        for (uint32_t j = a; j < b; j++) { src[j - w0] = (uint16_t)(pos - dist + r); r = r + 1u == dist ? 0u : r + 1u; }
    }
}
/* one pass over the lane's cells (lane, lane + 32, ...), eight cells at a time: all loads of a batch before its stores, so a pass
 * costs a few shared-memory latencies, not one per cell.  A literal looks itself up and stays; a source before the window is
 * final.  true = something moved */
ITX_HD bool itx_lzw_jump(uint16_t *src, uint32_t w0, uint32_t cnt, uint32_t lane) {
    bool changed = false;
    const uint32_t cnt_r = (cnt + 255u) & ~255u;
    for (uint32_t i0 = lane; i0 < cnt_r; i0 += 256u) {
        uint32_t sv[8], ss[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; u++) sv[u] = (uint32_t)src[i0 + 32u * u];
#pragma unroll
        for (uint32_t u = 0; u < 8u; u++) ss[u] = sv[u] >= w0 ? (uint32_t)src[sv[u] - w0] : sv[u];
#pragma unroll
        for (uint32_t u = 0; u < 8u; u++) if (ss[u] != sv[u]) { src[i0 + 32u * u] = (uint16_t)ss[u]; changed = true; }
    }
    return changed;
}
#endif
