"""The bigWig writer (iteres_b200/csrc/itx_bigwig.c, host C, the product's bigWigFileCreate) against the reference:
the committed golden .bigWig files of every `stat` / `cpgstat` known-answer variant are covered by the golden
suites; here a hg19-shaped world (1395 subfamilies: several sections per consensus, multi-level R and B+ trees,
several zoom levels) is compared with what the UNMODIFIED reference binary writes, where that binary exists
(oracle/_ref/iteres: the build container and the GPU box), and the error paths are checked."""
import ctypes as C
import filecmp
import os
import subprocess

import pytest

import synth
from iteres_b200 import capi

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "iteres")


def to_bigwig(wig, sizes, out):
    err = C.create_string_buffer(capi.ERRLEN)
    rc = capi.lib().itx_wig_to_bigwig(wig.encode(), sizes.encode(), out.encode(), err)
    return rc, err.value.decode()


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference binary is not built here")
@pytest.mark.parametrize("shape,n_rmsk,mode,n_units", [(1, 60000, 0, 150000), (0, 20000, 2, 30000)])
def test_same_bytes_as_the_reference_binary(shape, n_rmsk, mode, n_units, tmp_path):
    d = str(tmp_path)
    s = synth.Synth(shape, n_rmsk, seed=11)
    cs, rs, rm = s.write_tables(d)
    bam = os.path.join(d, "reads.bam")
    s.write_bam(bam, mode, n_units, level=1, threads=4)
    s.close()
    p = subprocess.run([REF, "stat", "-w", "-o", "ref", cs, rs, rm, bam], cwd=d, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-500:]
    for stem in ("ref.iteres", "ref.iteres.unique"):
        mine = os.path.join(d, stem + ".mine.bigWig")
        assert to_bigwig(os.path.join(d, stem + ".wig"), rs, mine) == (0, "")
        assert filecmp.cmp(mine, os.path.join(d, stem + ".bigWig"), shallow=False), stem


def test_errors_follow_the_reference(tmp_path):
    sizes = tmp_path / "rep.sizes"
    sizes.write_text("AluY\t300\nL1\t10\nAluY\t5\n")               # a later row of the same name wins
    wig = tmp_path / "x.wig"
    wig.write_text("")
    rc, msg = to_bigwig(str(wig), str(sizes), str(tmp_path / "x.bigWig"))
    assert rc == -2 and msg.endswith("is empty of data")         # bwgCreate.c:1108-1109
    wig.write_text("fixedStep chrom=L1 start=1 step=1 span=1\n" + "1\n" * 11)
    rc, msg = to_bigwig(str(wig), str(sizes), str(tmp_path / "x.bigWig"))
    assert rc == -2 and "has 10 bases, but item ends at 11" in msg
    wig.write_text("fixedStep chrom=AluY start=1 step=1 span=1\n" + "1\n" * 6)
    rc, msg = to_bigwig(str(wig), str(sizes), str(tmp_path / "x.bigWig"))
    assert rc == -2 and "has 5 bases" in msg
    wig.write_text("fixedStep chrom=MIR start=1 step=1 span=1\n1\n")
    rc, msg = to_bigwig(str(wig), str(sizes), str(tmp_path / "x.bigWig"))
    assert rc == -2 and "'MIR' not found" in msg
    wig.write_text("fixedStep chrom=L1 start=1 step=1 span=1\n" + "2\n" * 10)
    assert to_bigwig(str(wig), str(sizes), str(tmp_path / "x.bigWig")) == (0, "")
    assert to_bigwig(str(tmp_path / "nope.wig"), str(sizes), str(tmp_path / "x.bigWig"))[0] == -1
