"""ctypes binding of the deterministic synthetic-input generator (tools/libitx_synth.so): chrom/repeat
size files, rmsk.txt, uncompressed BAM record streams (any chunk range, for per-rank shards), BGZF
.bam files and CpG bedGraphs of the shapes SURVEY.md 8(d) / BASELINE.json name."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "tools", "libitx_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, cp, u64 = C.c_void_p, C.c_char_p, C.c_uint64
        L.synth_new.restype = vp
        L.synth_new.argtypes = [C.c_int, u64, C.c_int, C.c_int, C.c_int, u64]
        L.synth_free.argtypes = [vp]
        L.synth_n_rmsk.restype = u64
        L.synth_n_rmsk.argtypes = [vp]
        L.synth_write_sizes.argtypes = [vp, cp, cp]
        L.synth_write_rmsk.argtypes = [vp, cp]
        L.synth_bam_header.restype = u64
        L.synth_bam_header.argtypes = [vp, vp, u64]
        L.synth_n_chunks.restype = u64
        L.synth_n_chunks.argtypes = [u64]
        L.synth_records_size.restype = u64
        L.synth_records_size.argtypes = [vp, C.c_int, u64, u64, u64, u64, C.c_int, C.POINTER(u64)]
        L.synth_records_fill.restype = u64
        L.synth_records_fill.argtypes = [vp, C.c_int, u64, u64, u64, u64, vp, C.c_int]
        L.synth_write_bam.argtypes = [cp, vp, u64, vp, u64, C.c_int, C.c_int]
        L.synth_write_bedgraph.argtypes = [vp, cp, u64, u64]
        _lib = L
    return _lib


class Synth:
    """shape 0: chr1 only (cfg1), shape 1: hg19 (cfg2..5).  mode 0 SE-50, 1 SE-75 + XA/NM, 2 PE-100."""

    def __init__(self, shape, n_rmsk, n_subfam=None, seed=1):
        self.shape, self.seed = shape, seed
        n_subfam = n_subfam or (1395 if shape else 1200)
        n_fam, n_cla = (56, 21) if shape else (60, 12)
        self.h = lib().synth_new(shape, n_rmsk, n_subfam, n_fam, n_cla, seed)
        self.n_rmsk = lib().synth_n_rmsk(self.h)

    def close(self):
        if self.h:
            lib().synth_free(self.h)
            self.h = None

    def write_tables(self, outdir):
        os.makedirs(outdir, exist_ok=True)
        p = lambda n: os.path.join(outdir, n)
        assert lib().synth_write_sizes(self.h, p("chrom.sizes").encode(), p("rep.sizes").encode()) == 0
        assert lib().synth_write_rmsk(self.h, p("rmsk.txt").encode()) == 0
        return p("chrom.sizes"), p("rep.sizes"), p("rmsk.txt")

    def header(self):
        n = lib().synth_bam_header(self.h, None, 0)
        buf = np.zeros(n, dtype=np.uint8)
        assert lib().synth_bam_header(self.h, buf.ctypes.data, n) == n
        return buf

    def n_chunks(self, n_units):
        return lib().synth_n_chunks(n_units)

    def records_size(self, mode, n_units, c0=0, c1=None, threads=8):
        c1 = self.n_chunks(n_units) if c1 is None else c1
        nrec = C.c_uint64(0)
        sz = lib().synth_records_size(self.h, mode, n_units, self.seed, c0, c1, threads, C.byref(nrec))
        return sz, nrec.value

    def records_into(self, ptr, mode, n_units, c0=0, c1=None, threads=8):
        c1 = self.n_chunks(n_units) if c1 is None else c1
        return lib().synth_records_fill(self.h, mode, n_units, self.seed, c0, c1, ptr, threads)

    def stream(self, mode, n_units, c0=0, c1=None, threads=8, slack=64):
        """header + records of chunks [c0,c1) as one uint8 array (with `slack` zero bytes after the end).
        Returns (array, n_bytes, n_records)."""
        hdr = self.header()
        sz, nrec = self.records_size(mode, n_units, c0, c1, threads)
        a = np.zeros(len(hdr) + sz + slack, dtype=np.uint8)
        a[: len(hdr)] = hdr
        if sz:
            got = self.records_into(a.ctypes.data + len(hdr), mode, n_units, c0, c1, threads)
            assert got == sz
        return a, len(hdr) + sz, nrec

    def write_bam(self, path, mode, n_units, level=1, threads=8):
        a, n, nrec = self.stream(mode, n_units, threads=threads, slack=0)
        hl = len(self.header())
        rc = lib().synth_write_bam(path.encode(), a.ctypes.data, hl, a.ctypes.data + hl, n - hl, level, threads)
        assert rc == 0
        return n, nrec

    def write_bedgraph(self, path, n_rows):
        assert lib().synth_write_bedgraph(self.h, path.encode(), n_rows, self.seed) == 0
