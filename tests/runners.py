"""Drive the oracle (and, in test_parity_*.py, the CUDA product) the way the reference CLI drives its
hot path: same option letters, defaults and output file names as stat.c:46-111, filter.c:46-141,
cpgstat.c:27-59, cpgfilter.c:31-94."""
import getopt
import os

import oracle_lib as O


def parse(cmd, args):
    """-> dict of options in the reference's own vocabulary."""
    o = dict(cmd=cmd, Q=10, cov=1e-4, diff=1, N=0, U=0, R=0, T=0, D=0, C=0, E=150, I=500, t=1, r=0, field=0, name="ALL",
             thr=0.0, B=0, V=0, S=1 if "-S" in args else 0)
    if cmd == "stat":
        opts, _ = getopt.getopt(args, "SQ:c:xN:U:RTDwBVCo:E:I:h?")
        for k, v in opts:
            if k == "-Q": o["Q"] = int(v, 0)
            elif k == "-c": o["cov"] = float(v)
            elif k == "-x": o["diff"] = 0
            elif k == "-N": o["N"] = int(v, 0)
            elif k == "-U": o["U"] = int(v, 0)
            elif k == "-R": o["R"] = 1
            elif k == "-T": o["T"] = 1
            elif k == "-D": o["D"] = 1
            elif k == "-C": o["C"] = 1
            elif k == "-E": o["E"] = int(v, 0)
            elif k == "-I": o["I"] = int(v, 0)
            elif k == "-B": o["B"] = 1
            elif k == "-V": o["V"] = 1
        o["nindex"] = {0: 9, 1: 8, 2: 6, 3: 0}[o["N"]]
        o["nindex2"] = {0: 10, 1: 7, 2: 0}[o["U"]]
    elif cmd == "filter":
        o["diff"] = 0
        opts, _ = getopt.getopt(args, "SQ:g:N:n:c:t:f:rRTDCE:I:o:h?")
        for k, v in opts:
            if k == "-Q": o["Q"] = int(v, 0)
            elif k == "-g": o["cov"] = float(v)
            elif k == "-N": o["N"] = int(v, 0)
            elif k == "-t": o["t"] = int(v, 0)
            elif k == "-r": o["r"] = 1
            elif k == "-R": o["R"] = 1
            elif k == "-T": o["T"] = 1
            elif k == "-D": o["D"] = 1
            elif k == "-C": o["C"] = 1
            elif k == "-n": o["field"], o["name"] = 10, v
            elif k == "-c": o["field"], o["name"] = 11, v
            elif k == "-f": o["field"], o["name"] = 12, v
            elif k == "-E": o["E"] = int(v, 0)
            elif k == "-I": o["I"] = int(v, 0)
        o["nindex"] = {0: 7, 1: 8, 2: 6, 3: 4}[o["N"]]
    elif cmd == "cpgfilter":
        opts, _ = getopt.getopt(args, "n:c:f:t:o:h?")
        for k, v in opts:
            if k == "-n": o["field"], o["name"] = 10, v
            elif k == "-c": o["field"], o["name"] = 11, v
            elif k == "-f": o["field"], o["name"] = 12, v
            elif k == "-t": o["thr"] = float(v)
    if o["name"] == "ALL":
        o["field"] = 0
    return o


def ora_opts(o):
    return O.default_opts(mapQ=o["Q"], filter=1 if o["cmd"] == "filter" else 0, rmDup=o["R"], addChr=o["C"],
                          discardWrongEnd=o["D"], iSize=o["I"], extension=o["E"], minCoverage=o["cov"], treat=o["T"],
                          diffSubfam=o["diff"])


def run_oracle(inp, cmd, args, outdir, prefix="out"):
    """inp: directory holding chrom.sizes, rep.sizes, rmsk.txt, reads.bam / cpg.bedGraph."""
    o = parse(cmd, args)
    f = lambda n: os.path.join(inp, n)
    p = os.path.join(outdir, prefix)
    ix = O.OracleIndex(f("chrom.sizes"), f("rep.sizes"), f("rmsk.txt"), o["field"], o["name"])
    try:
        if cmd == "stat":
            ix.scan_file(f("reads.bam"), ora_opts(o))
            ix.write_stat(p, o["nindex"], o["nindex2"])
            ix.write_report(p + ".iteres.report", o["Q"], "ALL")
        elif cmd == "filter":
            ix.scan_file(f("reads.bam"), ora_opts(o))
            ix.write_filter("%s_%s.iteres.loci" % (p, o["name"]), o["r"], o["t"], o["nindex"])
            ix.write_report("%s_%s.iteres.reportloci" % (p, o["name"]), o["Q"], o["name"])
        elif cmd == "cpgstat":
            ix.scan_cpg(f("cpg.bedGraph"), 0)
            ix.write_cpg_stat(p)
        elif cmd == "cpgfilter":
            ix.scan_cpg(f("cpg.bedGraph"), 1)
            ix.write_cpg_filter("%s_%s.CpG.loci" % (p, o["name"]), o["thr"])
        return list(ix.cnt)
    finally:
        ix.close()


def itx_opts(o, prefix=None):
    from iteres_b200 import capi
    kw = {}
    if o["cmd"] == "stat" and prefix:                  # -B / -V: bed files next to the tables (stat.c:101-111)
        if o["B"]: kw["outbed"] = (prefix + ".iteres.bed").encode()
        if o["V"]: kw["outbed_unique"] = (prefix + ".iteres.unique.bed").encode()
    if o["cmd"] == "filter" and o["r"]:
        kw["readNames"] = 1
    if o["S"]:
        kw["isSam"] = 1
    return capi.default_opts(mapQ=o["Q"], filter=1 if o["cmd"] == "filter" else 0, rmDup=o["R"], addChr=o["C"],
                             discardWrongEnd=o["D"], iSize=o["I"], extension=o["E"], minCoverage=o["cov"], treat=o["T"],
                             diffSubfam=o["diff"], **kw)


def run_itx(make_index, scan, inp, cmd, args, outdir, prefix="out"):
    """Same driver for the CUDA product (iteres_b200.Index) and the host emulation of its device logic
    (emu_lib.EmuIndex).  make_index(chrom, rep, rmsk, field, name) -> index; scan(ix, bam_path, opts)."""
    o = parse(cmd, args)
    f = lambda n: os.path.join(inp, n)
    p = os.path.join(outdir, prefix)
    ix = make_index(f("chrom.sizes"), f("rep.sizes"), f("rmsk.txt"), o["field"], o["name"])
    reads = f("reads.sam") if o["S"] else f("reads.bam")
    try:
        if cmd == "stat":
            scan(ix, reads, itx_opts(o, p))
            ix.write_stat(p, o["nindex"], o["nindex2"])
            wig_to_bigwig(p + ".iteres.wig", f("rep.sizes"), p + ".iteres.bigWig")
            wig_to_bigwig(p + ".iteres.unique.wig", f("rep.sizes"), p + ".iteres.unique.bigWig")
            ix.write_report(p + ".iteres.report", o["Q"], "ALL")
        elif cmd == "filter":
            scan(ix, reads, itx_opts(o))
            ix.write_filter("%s_%s.iteres.loci" % (p, o["name"]), o["r"], o["t"], o["nindex"])
            ix.write_report("%s_%s.iteres.reportloci" % (p, o["name"]), o["Q"], o["name"])
        elif cmd == "cpgstat":
            ix.scan_cpg(f("cpg.bedGraph"), 0)
            ix.write_cpg_stat(p)
            wig_to_bigwig(p + ".CpGstat.wig", f("rep.sizes"), p + ".CpGstat.bigWig")
        elif cmd == "cpgfilter":
            ix.scan_cpg(f("cpg.bedGraph"), 1)
            ix.write_cpg_filter("%s_%s.CpG.loci" % (p, o["name"]), o["thr"])
        return list(ix.cnt)
    finally:
        ix.close()


def sam_to_bam_bytes(path):
    """what -S scans: the product's SAM text front end (host C in libiteres_gpu.so; needs no device)"""
    import ctypes as C
    from iteres_b200 import capi
    L = capi.lib()
    err = C.create_string_buffer(capi.ERRLEN)
    p, n = C.c_void_p(), C.c_uint64()
    rc = L.itx_sam_to_bam(path.encode(), C.byref(p), C.byref(n), err)
    if rc:
        raise capi.ItxError(rc, err.value.decode())
    data = C.string_at(p, n.value)
    L.itx_free(p)
    return data


def wig_to_bigwig(wig, sizes, out):
    """the product's bigWigFileCreate (host C in libiteres_gpu.so; needs no device)"""
    import ctypes as C
    from iteres_b200 import capi
    err = C.create_string_buffer(capi.ERRLEN)
    rc = capi.lib().itx_wig_to_bigwig(wig.encode(), sizes.encode(), out.encode(), err)
    if rc:
        raise capi.ItxError(rc, err.value.decode())


def needs_host_order(cmd, args):
    """every variant is produced now: -R on the device, bed lines and filter -r name lists by the host pass over the
    device's per-record verdicts"""
    return False


SKIP = {"cmdline.txt", "stderr.txt"}


def expected_files(vdir, skip_bed=False):
    """skip_bed: the oracle restates the counting path only -- no bed lines, no bigWig container"""
    out = []
    for fn in sorted(os.listdir(vdir)):
        if fn in SKIP or (skip_bed and (fn.endswith(".bed") or fn.endswith(".bigWig"))):
            continue
        out.append(fn)
    return out
