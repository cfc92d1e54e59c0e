"""Two GPUs of one box, one index per device: every rank scans its record-range shard, ONE NCCL allreduce of the packed
integer counter block, and rank 0's tables are the single-GPU (= the oracle's) tables, bit for bit -- integer sums do
not depend on who adds them.  Needs two devices (`gpurun --gpus 2`); skipped otherwise."""
import threading

import numpy as np
import pytest

import oracle_lib as O
import synth
from iteres_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("filt", [0, 1], ids=["stat", "filter"])
def test_two_gpu_shards_equal_the_whole(filt, tmp_path):
    if capi.lib().itx_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=13)
    cs, rs, rm = s.write_tables(d)
    n_units, mode, world = 120000, 2, 2
    nch = s.n_chunks(n_units)
    whole, n, nrec = s.stream(mode, n_units)
    ora = O.OracleIndex(cs, rs, rm)
    want = ora.scan_stream(whole[:n].tobytes(), O.default_opts(filter=filt, diffSubfam=0 if filt else 1))
    shards = [s.stream(mode, n_units, r * nch // world, (r + 1) * nch // world) for r in range(world)]
    ixs = [capi.Index(cs, rs, rm, device=r) for r in range(world)]
    uid = ixs[0].comm_unique_id()
    errs = []

    def rank(r):
        try:
            ixs[r].comm_init(uid, r, world)
            buf, nb, _ = shards[r]
            ixs[r].scan_bam_host(buf.ctypes.data, nb, capi.default_opts(filter=filt, diffSubfam=0 if filt else 1))
            ixs[r].allreduce_counts()
        except Exception as e:           # noqa: BLE001
            errs.append((r, e))
    th = [threading.Thread(target=rank, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for r in range(world):
        assert list(ixs[r].cnt) == want, r
    L = O.lib()
    import ctypes as C
    c4 = (C.c_uint64 * 4)()
    if filt == 0:
        for which, nfun in ((0, L.ora_n_subfam), (1, L.ora_n_fam), (2, L.ora_n_class)):
            got = ixs[0].table(which)
            assert len(got) == nfun(ora.h)
            for i, row in enumerate(got):
                L.ora_counts(ora.h, which, i, c4)
                assert row == (L.ora_name(ora.h, which, i).decode(),) + tuple(c4)
        for i in range(0, L.ora_n_subfam(ora.h), 7):
            ln = L.ora_subfam_length(ora.h, i)
            if ln:
                assert np.array_equal(ixs[0].coverage(i, 0), np.ctypeslib.as_array(L.ora_subfam_bp(ora.h, i, 0), shape=(ln,)))
    else:
        got = ixs[0].elem_counts_by_row(0)
        assert int(got.sum()) == want[9] and want[9] > 0
    for ix in ixs:
        ix.close()
    ora.close()
    s.close()


def _two_ranks(ixs, fn):
    errs = []

    def rank(r):
        try:
            fn(r)
        except Exception as e:           # noqa: BLE001
            errs.append((r, e))
    th = [threading.Thread(target=rank, args=(r,)) for r in range(len(ixs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


@pytest.mark.parametrize("mode", [0, 2], ids=["se50", "pe100"])
def test_two_gpus_one_file(mode, tmp_path):
    """ONE BGZF file, two devices: itx_scan_alignments_shard (block ranges, guessed first records, NCCL all-gather of the reports,
    chain check) + ONE allreduce = the single-device scan of the file, bit for bit"""
    if capi.lib().itx_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import os
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=14)
    cs, rs, rm = s.write_tables(d)
    bam = os.path.join(d, "reads.bam")
    s.write_bam(bam, mode, 200000, level=1, threads=4)
    whole = capi.Index(cs, rs, rm, device=0)
    want = whole.scan_alignments(bam, capi.default_opts())
    world = 2
    ixs = [capi.Index(cs, rs, rm, device=r) for r in range(world)]
    uid = ixs[0].comm_unique_id()

    def run(r):
        ixs[r].comm_init(uid, r, world)
        ixs[r].scan_alignments_shard(bam, capi.default_opts())
        ixs[r].allreduce_counts()
    _two_ranks(ixs, run)
    for r in range(world):
        assert list(ixs[r].cnt) == want, r
    for which in range(3):
        assert ixs[0].table(which) == whole.table(which)
    for i in range(0, whole.n(0), 5):
        for u in (0, 1):
            assert np.array_equal(ixs[0].coverage(i, u), whole.coverage(i, u))
    for ix in ixs + [whole]:
        ix.close()
    s.close()


def test_two_gpus_cpg_parts(tmp_path):
    """cpgstat: each device takes its part of the bedGraph's lines; u32 counts and f64 score sums go through the allreduce"""
    if capi.lib().itx_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import os
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=15)
    cs, rs, rm = s.write_tables(d)
    bg = os.path.join(d, "cpg.bedGraph")
    s.write_bedgraph(bg, 200000)
    whole = capi.Index(cs, rs, rm, device=0)
    lines, inrep = whole.scan_cpg(bg)
    world = 2
    ixs = [capi.Index(cs, rs, rm, device=r) for r in range(world)]
    uid = ixs[0].comm_unique_id()

    def run(r):
        ixs[r].comm_init(uid, r, world)
        ixs[r].scan_cpg_shard(bg, r, world)
        ixs[r].allreduce_counts()
    _two_ranks(ixs, run)
    assert ixs[0].cpg_totals() == (lines, inrep) == ixs[1].cpg_totals()
    a, b = os.path.join(d, "two"), os.path.join(d, "one")
    ixs[0]._dirty = True
    ixs[0].write_cpg_stat(a)
    whole.write_cpg_stat(b)
    from test_gpu_parity import close_text
    for nme in (".CpG.subfamily.stat", ".CpGstat.wig", ".CpG.family.stat", ".CpG.class.stat"):
        assert close_text(a + nme, b + nme), nme
    for ix in ixs + [whole]:
        ix.close()
    s.close()


def test_command_line_on_two_gpus(tmp_path):
    """ITERES_GPUS=2 iteres stat / filter / cpgstat write the files ITERES_GPUS=1 writes"""
    if capi.lib().itx_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import filecmp
    import os
    import subprocess
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=16)
    cs, rs, rm = s.write_tables(d)
    bam, bg = os.path.join(d, "reads.bam"), os.path.join(d, "cpg.bedGraph")
    s.write_bam(bam, 2, 150000, level=1, threads=4)
    s.write_bedgraph(bg, 100000)
    exe = os.path.join(os.path.dirname(capi.lib_path()), "iteres")
    from test_gpu_parity import close_text
    for cmd, inp in (("stat", bam), ("filter", bam), ("cpgstat", bg)):
        outs = []
        for n in (1, 2):
            od = os.path.join(d, "%s_%d" % (cmd, n))
            os.makedirs(od)
            p = subprocess.run([exe, cmd, "-o", "out", cs, rs, rm, inp], cwd=od, env=dict(os.environ, ITERES_GPUS=str(n)), capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
            outs.append(od)
        names = sorted(os.listdir(outs[0]))
        assert names == sorted(os.listdir(outs[1])) and names
        for fn in names:
            a, b = os.path.join(outs[0], fn), os.path.join(outs[1], fn)
            assert filecmp.cmp(a, b, shallow=False) or (cmd == "cpgstat" and not fn.endswith(".bigWig") and close_text(a, b)), (cmd, fn)
    s.close()
