/* itx_inflate.cuh -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for one GPU thread per block.
 *
 * BGZF (cussamtools/bgzf.c:367-397 inflate_block: inflateInit2(-15) + inflate(Z_FINISH), no CRC check)
 * stores every <= 64 KiB of the BAM stream as an independent raw-deflate stream, so the blocks of a file
 * can be inflated by tens of thousands of threads at once.  The decoder is canonical-Huffman, bit-serial
 * (count[] / symbol[] per code, no lookup tables) so that a thread's whole state is ~1 KiB and can live in
 * shared memory, interleaved across the threads of a CTA to stay clear of bank conflicts.
 *
 * `Tab` is the per-thread table store: tab(j) is the j-th 16-bit cell of this thread.  On the device it
 * is shared memory strided by the CTA size; the test-only host build uses a plain array.
 * Like itx_logic.cuh this file is __host__ __device__ so the non-GPU suite can check it against zlib.
 */
#ifndef ITX_INFLATE_CUH
#define ITX_INFLATE_CUH
#include "itx_logic.cuh"

#define ITX_INF_OK 0
#define ITX_INF_EDATA 1        /* invalid deflate data */
#define ITX_INF_ESIZE 2        /* output is not the ISIZE the BGZF footer promised */

/* cell layout of one thread's table store (16-bit cells) */
#define ITX_T_LCNT 0           /* [16]  literal/length code: number of codes of each length */
#define ITX_T_LSYM 16          /* [288] literal/length symbols in canonical order */
#define ITX_T_DCNT 304         /* [16]  distance code counts */
#define ITX_T_DSYM 320         /* [30]  distance symbols */
#define ITX_T_LENS 350         /* [160] code lengths while a dynamic header is read, two 8-bit lengths per cell */
#define ITX_T_CELLS 510

template <class Tab>
struct itx_inflater {
    const uint8_t *in; uint32_t in_len, in_pos;     /* compressed bytes (over-reads of up to 8 bytes are harmless: buffers carry slack) */
    uint8_t *out; uint32_t out_cap, out_pos;
    uint64_t bitbuf; uint32_t bitcnt;
    Tab tab;
    uint32_t err;

    /* four bytes at a time; reads may run a few bytes past the stream (buffers carry slack), the bits are never used */
    ITX_HDM void refill() {
        if (bitcnt <= 32) {
            const uint8_t *q = in + in_pos;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3) * 8;
            const uint32_t v = sh ? itx_funnel_r(w[0], w[1], sh) : w[0];
            bitbuf |= (uint64_t)v << bitcnt; bitcnt += 32; in_pos += 4;
        }
    }
    ITX_HDM uint32_t bits(uint32_t n) {                 /* n <= 16 */
        refill();
        const uint32_t v = (uint32_t)bitbuf & ((1u << n) - 1u);
        bitbuf >>= n; bitcnt -= n;
        return v;
    }
    ITX_HDM uint32_t get_len(uint32_t i) const { const uint32_t c = tab(ITX_T_LENS + (i >> 1)); return (i & 1) ? (c >> 8) : (c & 0xff); }
    ITX_HDM void set_len(uint32_t i, uint32_t v) {
        const uint32_t c = tab(ITX_T_LENS + (i >> 1));
        tab.set(ITX_T_LENS + (i >> 1), (uint16_t)((i & 1) ? ((c & 0x00ff) | (v << 8)) : ((c & 0xff00) | v)));
    }
    /* canonical Huffman decode, one bit at a time.  The 15 per-length code counts of a table sit in eight
     * registers (two 16-bit counts each, loaded by load_counts after a table is built), so the only memory
     * access of a decode is the final symbol lookup. */
    uint32_t lc[8], dc[8];
    ITX_HDM void load_counts(uint32_t cnt, uint32_t c[8]) {
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) c[k] = (uint32_t)tab(cnt + 2 * k) | ((uint32_t)tab(cnt + 2 * k + 1) << 16);
    }
    ITX_HDM int32_t decode_regs(const uint32_t c[8], uint32_t sym) {
        refill();
        int32_t code = 0, first = 0, index = 0;
        uint32_t buf = (uint32_t)bitbuf;
#pragma unroll
        for (uint32_t len = 1; len <= 15; len++) {
            code |= (int32_t)(buf & 1u); buf >>= 1;
            const int32_t count = (int32_t)((len & 1) ? (c[len >> 1] >> 16) : (c[len >> 1] & 0xffffu));
            if (code - count < first) { bitbuf >>= len; bitcnt -= len; return (int32_t)tab(sym + (uint32_t)(index + (code - first))); }
            index += count; first += count; first <<= 1; code <<= 1;
        }
        return -1;
    }
    /* the same with the counts read from the table store (code-length code of a dynamic header) */
    ITX_HDM int32_t decode(uint32_t cnt, uint32_t sym) {
        refill();
        int32_t code = 0, first = 0, index = 0;
        uint32_t buf = (uint32_t)bitbuf;
        for (uint32_t len = 1; len <= 15; len++) {
            code |= (int32_t)(buf & 1u); buf >>= 1;
            const int32_t count = (int32_t)tab(cnt + len);
            if (code - count < first) { bitbuf >>= len; bitcnt -= len; return (int32_t)tab(sym + (uint32_t)(index + (code - first))); }
            index += count; first += count; first <<= 1; code <<= 1;
        }
        return -1;
    }
    /* build count[] / symbol[] from lengths first..first+n-1 (read through len_at); returns 0 for a complete
     * code, <0 over-subscribed, >0 incomplete (allowed only for a single-code distance set, as zlib does) */
    template <class LenAt>
    ITX_HDM int32_t construct(uint32_t cnt, uint32_t sym, uint32_t n, LenAt len_at) {
        for (uint32_t l = 0; l <= 15; l++) tab.set(cnt + l, 0);
        for (uint32_t s = 0; s < n; s++) { const uint32_t l = len_at(s); tab.set(cnt + l, (uint16_t)(tab(cnt + l) + 1)); }
        if (tab(cnt) == n) return 0;                    /* no codes at all: complete, but decode() will fail */
        int32_t left = 1;
        for (uint32_t l = 1; l <= 15; l++) { left <<= 1; left -= (int32_t)tab(cnt + l); if (left < 0) return left; }
        /* offsets of each length in the symbol table, kept in 16 registers */
        uint16_t offs[16]; offs[1] = 0;
#pragma unroll
        for (uint32_t l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + tab(cnt + l));
        for (uint32_t s = 0; s < n; s++) {
            const uint32_t l = len_at(s);
            if (l) {
                uint16_t o = 0;
#pragma unroll
                for (uint32_t k = 1; k <= 15; k++) if (k == l) { o = offs[k]; offs[k] = (uint16_t)(o + 1); }
                tab.set(sym + o, (uint16_t)s);
            }
        }
        return left;
    }
    ITX_HDM void put(uint8_t b) { if (out_pos < out_cap) out[out_pos] = b; out_pos++; }

    /* length / distance base values and extra bits (RFC 1951 3.2.5), computed instead of tabulated */
    static ITX_HDM uint32_t lext(uint32_t s) { return (s < 8 || s == 28) ? 0u : (s - 4) >> 2; }
    static ITX_HDM uint32_t lbase(uint32_t s) { return s < 8 ? 3 + s : (s == 28 ? 258u : 3 + ((4 + (s & 3)) << ((s - 4) >> 2))); }
    static ITX_HDM uint32_t dext(uint32_t d) { return d < 4 ? 0u : (d - 2) >> 1; }
    static ITX_HDM uint32_t dbase(uint32_t d) { return d < 4 ? 1 + d : 1 + ((2 + (d & 1)) << ((d - 2) >> 1)); }
    /* one literal / length-distance pair / end-of-block; false = invalid data */
    ITX_HDM bool symbol(bool *end_of_block) {
        *end_of_block = false;
        if (out_pos > out_cap || in_pos > in_len + 8) return false;
        int32_t s = decode_regs(lc, ITX_T_LSYM);
        if (s < 0) return false;
        if (s < 256) { put((uint8_t)s); return true; }
        if (s == 256) { *end_of_block = true; return true; }
        s -= 257;
        if (s >= 29) return false;
        const uint32_t len = lbase((uint32_t)s) + bits(lext((uint32_t)s));
        const int32_t d = decode_regs(dc, ITX_T_DSYM);
        if (d < 0 || d >= 30) return false;
        const uint32_t dist = dbase((uint32_t)d) + bits(dext((uint32_t)d));
        if (dist > out_pos) return false;
        if (out_pos + len > out_cap) { out_pos += len; return false; }
        uint8_t *o = out + out_pos; const uint8_t *f = o - dist;
        uint32_t k = 0;
        if (dist >= 36) {
            /* 32 bytes per round out of nine aligned words: the loads of a round are independent of each other and
             * of the round's stores (the source lies at least 36 bytes behind), so one memory latency moves 32 bytes */
            for (; k + 32 <= len; k += 32) {
                const uint8_t *q = f + k;
                const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3) * 8;
                uint32_t v[9];
#pragma unroll
                for (int j = 0; j < 9; j++) v[j] = w[j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t x = sh ? itx_funnel_r(v[j], v[j + 1], sh) : v[j];
                    o[k + 4 * j] = (uint8_t)x; o[k + 4 * j + 1] = (uint8_t)(x >> 8); o[k + 4 * j + 2] = (uint8_t)(x >> 16); o[k + 4 * j + 3] = (uint8_t)(x >> 24);
                }
            }
        }
        if (dist >= 8) {                                             /* eight independent loads in flight per round */
            for (; k + 8 <= len; k += 8) {
                const uint8_t t0 = f[k], t1 = f[k + 1], t2 = f[k + 2], t3 = f[k + 3], t4 = f[k + 4], t5 = f[k + 5], t6 = f[k + 6], t7 = f[k + 7];
                o[k] = t0; o[k + 1] = t1; o[k + 2] = t2; o[k + 3] = t3; o[k + 4] = t4; o[k + 5] = t5; o[k + 6] = t6; o[k + 7] = t7;
            }
        }
        for (; k < len; k++) o[k] = f[k];                            /* byte-wise: overlapping copies replicate, as LZ77 requires */
        out_pos += len;
        return true;
    }
    ITX_HDM bool stored() {
        bitbuf >>= (bitcnt & 7); bitcnt -= (bitcnt & 7);          /* to the next byte boundary */
        const uint32_t len = bits(16), nlen = bits(16);
        if ((len ^ 0xffffu) != nlen) return false;
        for (uint32_t k = 0; k < len; k++) put((uint8_t)bits(8));
        return in_pos <= in_len + 8;
    }
    ITX_HDM bool fixed() {
        for (uint32_t s = 0; s < 288; s++) set_len(s, s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
        itx_inflater *self = this;
        construct(ITX_T_LCNT, ITX_T_LSYM, 288, [self](uint32_t s) { return self->get_len(s); });
        for (uint32_t s = 0; s < 30; s++) set_len(s, 5);
        construct(ITX_T_DCNT, ITX_T_DSYM, 30, [self](uint32_t s) { return self->get_len(s); });
        load_counts(ITX_T_LCNT, lc); load_counts(ITX_T_DCNT, dc);
        return true;
    }
    ITX_HDM bool dynamic() {
        /* order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15, five bits each */
        const uint64_t O0 = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
        const uint64_t O1 = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
        const uint32_t nlen = bits(5) + 257, ndist = bits(5) + 1, ncode = bits(4) + 4;
        if (nlen > 286 || ndist > 30) return false;
        /* the code-length code: its 19 lengths sit at the front of the length area, its tables in the distance cells */
        for (uint32_t i = 0; i < 19; i++) set_len(300 + i, 0);
        for (uint32_t i = 0; i < ncode; i++) set_len(300 + (uint32_t)((i < 12 ? O0 >> (5 * i) : O1 >> (5 * (i - 12))) & 31), bits(3));
        itx_inflater *self = this;
        if (construct(ITX_T_DCNT, ITX_T_DSYM, 19, [self](uint32_t s) { return self->get_len(300 + s); }) != 0) return false;
        uint32_t i = 0;
        while (i < nlen + ndist) {
            int32_t s = decode(ITX_T_DCNT, ITX_T_DSYM);
            if (s < 0) return false;
            if (s < 16) set_len(i++, (uint32_t)s);
            else {
                uint32_t v = 0, rep;
                if (s == 16) { if (i == 0) return false; v = get_len(i - 1); rep = 3 + bits(2); }
                else if (s == 17) rep = 3 + bits(3);
                else rep = 11 + bits(7);
                if (i + rep > nlen + ndist) return false;
                while (rep--) set_len(i++, v);
            }
        }
        if (get_len(256) == 0) return false;                       /* no end-of-block code */
        /* the distance lengths follow the literal/length ones: move them out of the way of nothing -- they are read in place */
        int32_t e = construct(ITX_T_LCNT, ITX_T_LSYM, nlen, [self](uint32_t s) { return self->get_len(s); });
        if (e != 0 && (e < 0 || nlen != (uint32_t)(tab(ITX_T_LCNT) + tab(ITX_T_LCNT + 1)))) return false;
        e = construct(ITX_T_DCNT, ITX_T_DSYM, ndist, [self, nlen](uint32_t s) { return self->get_len(nlen + s); });
        if (e != 0 && (e < 0 || ndist != (uint32_t)(tab(ITX_T_DCNT) + tab(ITX_T_DCNT + 1)))) return false;
        load_counts(ITX_T_LCNT, lc); load_counts(ITX_T_DCNT, dc);
        return true;
    }
    /* The decoder is a small state machine so that the 32 lanes of a warp (32 different BGZF blocks) can be
     * stepped together: every call of advance() does ONE unit of work -- a block header (with its table build)
     * or one symbol -- and the kernel re-converges the warp between calls. */
    uint32_t state, last, expect;                /* state: 0 header, 1 symbols, 2 done, 3 error */
    ITX_HDM void begin(uint32_t expect_) { bitbuf = 0; bitcnt = 0; in_pos = 0; out_pos = 0; state = 0; last = 0; expect = expect_; err = ITX_INF_OK; }
    ITX_HDM void advance() {
        if (state == 1) {
            bool eob;
            if (!symbol(&eob)) { state = 3; err = ITX_INF_EDATA; }
            else if (eob) state = last ? 2u : 0u;
        } else if (state == 0) {
            last = bits(1);
            const uint32_t type = bits(2);
            bool ok;
            if (type == 0) { ok = stored(); state = last ? 2u : 0u; }
            else { ok = type == 1 ? fixed() : (type == 2 ? dynamic() : false); state = 1; }
            if (!ok || in_pos > in_len + 8) { state = 3; err = ITX_INF_EDATA; }
        }
        if (state == 2 && out_pos != expect) { state = 3; err = ITX_INF_ESIZE; }
    }
    /* one whole raw-deflate stream; returns ITX_INF_* */
    ITX_HDM uint32_t run(uint32_t expect_) {
        begin(expect_);
        while (state < 2) advance();
        return err;
    }
};
#endif
