"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "iteres")


class OraOpts(C.Structure):
    _fields_ = [("mapQ", C.c_uint32), ("filter", C.c_int32), ("rmDup", C.c_int32), ("addChr", C.c_int32),
                ("discardWrongEnd", C.c_int32), ("iSize", C.c_uint32), ("extension", C.c_uint32),
                ("minCoverage", C.c_float), ("treat", C.c_int32), ("diffSubfam", C.c_int32)]


class OraTrace(C.Structure):
    _fields_ = [("start", C.c_uint32), ("end", C.c_uint32), ("tid", C.c_int32), ("sel_row", C.c_int32),
                ("flags", C.c_uint32)]


def default_opts(**kw):
    o = OraOpts(10, 0, 0, 0, 0, 500, 150, 1e-4, 0, 1)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ora_index_build.restype = C.c_void_p
        L.ora_index_build.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        L.ora_index_free.argtypes = [C.c_void_p]
        L.ora_index_reset_counts.argtypes = [C.c_void_p]
        L.ora_scan_bam_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(OraOpts),
                                          C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ora_scan_bam_file.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(OraOpts), C.POINTER(C.c_uint64)]
        L.ora_inflate_bam.restype = C.c_void_p
        L.ora_inflate_bam.argtypes = [C.c_char_p, C.POINTER(C.c_uint64)]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_scan_cpg.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_char_p]
        L.ora_write_stat.argtypes = [C.c_void_p] + [C.c_char_p] * 5 + [C.c_uint64, C.c_uint64]
        L.ora_write_report.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_char_p]
        L.ora_write_filter.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_uint64]
        L.ora_write_cpg_stat.argtypes = [C.c_void_p] + [C.c_char_p] * 4
        L.ora_write_cpg_filter.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ora_find_select.restype = C.c_int32
        L.ora_find_select.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_float,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
        L.ora_n_subfam.argtypes = [C.c_void_p]
        L.ora_n_fam.argtypes = [C.c_void_p]
        L.ora_n_class.argtypes = [C.c_void_p]
        L.ora_n_elem.argtypes = [C.c_void_p]
        L.ora_n_elem.restype = C.c_int64
        L.ora_name.restype = C.c_char_p
        L.ora_name.argtypes = [C.c_void_p, C.c_int, C.c_int32]
        L.ora_counts.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.POINTER(C.c_uint64)]
        L.ora_subfam_length.restype = C.c_uint32
        L.ora_subfam_length.argtypes = [C.c_void_p, C.c_int32]
        L.ora_subfam_bp.restype = C.POINTER(C.c_uint32)
        L.ora_subfam_bp.argtypes = [C.c_void_p, C.c_int32, C.c_int]
        _lib = L
    return _lib


class OracleIndex:
    def __init__(self, chrom_sizes, rep_sizes, rmsk, filter_field=0, filter_name="ALL"):
        err = C.create_string_buffer(256)
        self.h = lib().ora_index_build(chrom_sizes.encode(), rep_sizes.encode(), rmsk.encode(), filter_field,
                                       filter_name.encode(), err)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.cnt = (C.c_uint64 * 13)()

    def close(self):
        if self.h:
            lib().ora_index_free(self.h)
            self.h = None

    def reset(self):
        lib().ora_index_reset_counts(self.h)

    def scan_file(self, path, opts):
        rc = lib().ora_scan_bam_file(self.h, path.encode(), C.byref(opts), self.cnt)
        if rc:
            raise RuntimeError("oracle scan failed")
        return list(self.cnt)

    def scan_stream(self, buf, opts, trace=False):
        """buf: bytes-like of the uncompressed BAM (header + records)."""
        import numpy as np
        a = np.frombuffer(buf, dtype=np.uint8)
        nrec = C.c_uint64(0)
        tr = None
        cap = 0
        if trace:
            cap = max(1, len(a) // 36)
            tr = (OraTrace * cap)()
        rc = lib().ora_scan_bam_stream(self.h, a.ctypes.data, len(a), C.byref(opts), self.cnt,
                                       C.cast(tr, C.c_void_p) if tr is not None else None, cap, C.byref(nrec))
        if rc:
            raise RuntimeError("oracle scan failed")
        if trace:
            t = np.ctypeslib.as_array(tr)[: nrec.value].copy() if nrec.value else np.zeros(0, dtype=np.dtype(OraTrace))
            return list(self.cnt), t
        return list(self.cnt)

    def write_stat(self, prefix, nindex=9, nindex2=10):
        f = lambda s: (prefix + s).encode()
        lib().ora_write_stat(self.h, f(".iteres.subfamily.stat"), f(".iteres.wig"), f(".iteres.family.stat"),
                             f(".iteres.class.stat"), f(".iteres.unique.wig"), self.cnt[nindex], self.cnt[nindex2])

    def write_report(self, path, mapQ=10, subfam="ALL"):
        lib().ora_write_report(path.encode(), self.cnt, mapQ, subfam.encode())

    def write_filter(self, path, readlist=0, threshold=1, nindex=7):
        lib().ora_write_filter(self.h, path.encode(), readlist, threshold, self.cnt[nindex])

    def scan_cpg(self, path, filter=0):
        a, b = C.c_uint32(0), C.c_uint32(0)
        err = C.create_string_buffer(256)
        rc = lib().ora_scan_cpg(self.h, path.encode(), filter, C.byref(a), C.byref(b), err)
        if rc:
            raise RuntimeError(err.value.decode())
        return a.value, b.value

    def write_cpg_stat(self, prefix):
        f = lambda s: (prefix + s).encode()
        lib().ora_write_cpg_stat(self.h, f(".CpG.subfamily.stat"), f(".CpGstat.wig"), f(".CpG.family.stat"), f(".CpG.class.stat"))

    def write_cpg_filter(self, path, thr=0.0):
        lib().ora_write_cpg_filter(self.h, path.encode(), thr)

    def find_select(self, chrom, start, end, min_cov=1e-4, cap=64):
        n = C.c_int32(0)
        hits = (C.c_int32 * cap)()
        sel = lib().ora_find_select(self.h, chrom.encode(), start, end, min_cov, C.byref(n), hits, cap)
        return sel, list(hits[: min(n.value, cap)])


def inflate_bam(path):
    n = C.c_uint64(0)
    p = lib().ora_inflate_bam(path.encode(), C.byref(n))
    if not p:
        raise RuntimeError("cannot read " + path)
    try:
        return C.string_at(p, n.value)
    finally:
        lib().ora_free(p)
