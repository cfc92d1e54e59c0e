"""ONE BGZF file split across ranks by BGZF block ranges (itx_scan_shard_file + itx_shard_chain_check): every
rank finds its first block by itself, guesses its first record start out of the bytes, reads past its last block for the record
that straddles it -- and the counts of all parts together are those of one pass over the file, bit for bit.  The ranks are run one
after the other on ONE device here (each on an index of its own; no communicator is needed for the protocol itself), which is the
same code path as N processes on N devices minus the NCCL all-gather; tests/test_multigpu.py runs that on two devices."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import bamio
import kats
import oracle_lib as O
import synth
from iteres_b200 import capi
from test_gpu_parity import assert_same_tables

pytestmark = pytest.mark.gpu


def run_parts(tabs, path, nranks, opts, filt=0, max_rounds=None):
    """the protocol of itx_scan_alignments_shard, the ranks taken in turn: scan every part from a guessed entry, check the chain,
    scan again where the check says so.  -> (indexes, reports, ranks that had to scan again)"""
    ixs = [capi.Index(*tabs) for _ in range(nranks)]
    reps, redone = [], []
    for r in range(nranks):
        _, rep = ixs[r].scan_shard_file(path, opts, r, nranks)
        reps.append(rep)
    for _ in range(max_rounds or nranks + 1):
        bad, forced = capi.shard_chain_check(reps)
        if bad < 0:
            break
        ixs[bad].reset()
        _, reps[bad] = ixs[bad].scan_shard_file(path, opts, bad, nranks, forced)
        redone.append(bad)
    else:
        raise AssertionError("the parts do not chain: %r" % (reps,))
    return ixs, reps, redone


def summed(ixs):
    cnt = [sum(int(ix.cnt[k]) for ix in ixs) for k in range(13)]
    tables = []
    for which in range(3):
        rows = [ix.table(which) for ix in ixs]
        tables.append([(rows[0][i][0],) + tuple(sum(r[i][k] for r in rows) for k in (1, 2)) + rows[0][i][3:] for i in range(len(rows[0]))])
    return cnt, tables


def assert_parts_equal_whole(ixs, whole):
    cnt, tables = summed(ixs)
    assert cnt == list(whole.cnt)
    for which in range(3):
        assert tables[which] == whole.table(which)
    for i in range(0, whole.n(0), 3):
        for u in (0, 1):
            want = whole.coverage(i, u)
            if len(want):
                got = sum((ix.coverage(i, u).astype(np.uint64) for ix in ixs), np.zeros(len(want), dtype=np.uint64))
                assert np.array_equal((got & 0xffffffff).astype(np.uint32), want), (i, u)


@pytest.mark.parametrize("nranks", [2, 3, 5, 8])
@pytest.mark.parametrize("mode", [0, 2], ids=["se50", "pe100"])
def test_block_range_parts_sum_to_the_whole(mode, nranks, tmp_path):
    """part boundaries fall wherever the byte count says: in the middle of records, of pairs, of BGZF blocks"""
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=21)
    tabs = s.write_tables(d)
    bam = os.path.join(d, "reads.bam")
    n, nrec = s.write_bam(bam, mode, 150000, level=1, threads=4)
    opts = capi.default_opts()
    whole = capi.Index(*tabs)
    want = whole.scan_alignments(bam, opts)
    ora = O.OracleIndex(*tabs)
    assert ora.scan_file(bam, O.default_opts()) == want and want[0] + want[1] == nrec
    ixs, reps, redone = run_parts(tabs, bam, nranks, opts)
    assert sum(r[2] for r in reps) == n                       # the parts tile the uncompressed stream
    assert reps[0][0] < reps[0][2] and all(r[1] < 1 << 20 for r in reps[:-1])
    assert reps[-1][1] == 0                                   # the last part's chain ends exactly at the end of the stream
    assert_parts_equal_whole(ixs, whole)
    # the whole-file tables are the oracle's
    assert_same_tables(whole, ora)
    for ix in ixs + [whole]:
        ix.close()
    ora.close()
    s.close()


def test_filter_mode_parts(tmp_path):
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=22)
    tabs = s.write_tables(d)
    bam = os.path.join(d, "reads.bam")
    s.write_bam(bam, 1, 120000, level=6, threads=4)
    opts = capi.default_opts(filter=1, diffSubfam=0)
    whole = capi.Index(*tabs)
    want = whole.scan_alignments(bam, opts)
    ixs, reps, _ = run_parts(tabs, bam, 4, opts)
    assert [sum(int(ix.cnt[k]) for ix in ixs) for k in range(13)] == want
    got = sum((ix.elem_counts_by_row(0).astype(np.uint64) for ix in ixs), np.zeros(len(whole.elem_counts_by_row(0)), dtype=np.uint64))
    assert np.array_equal(got.astype(np.uint32), whole.elem_counts_by_row(0)) and int(got.sum()) == want[9] > 0
    for ix in ixs + [whole]:
        ix.close()
    s.close()


def _kat_tables(d):
    cs, rs, rm = (os.path.join(d, n) for n in ("chrom.sizes", "rep.sizes", "rmsk.txt"))
    open(cs, "w").write("chr1\t1000000\n")
    open(rs, "w").write("AluY\t300\nL1PA2\t6000\n")
    open(rm, "w").write("\n".join(kats.ANNOT1) + "\n")
    return cs, rs, rm


@pytest.mark.parametrize("payload_kb", [200, 3000], ids=["200KB_record", "3MB_record_margin_retry"])
def test_a_wrong_first_record_guess_is_caught_between_ranks(payload_kb, tmp_path):
    """the second rank's part begins inside a huge byte-array aux field full of back-to-back copies of a valid record: it
    guesses one of the copies as its first record, the chain check against the first rank's exit says otherwise, and the part
    is scanned again from the true record start.  With a 3 MB record the first rank also has to come back for more bytes
    (its 1 MiB margin does not hold the record that straddles its end)."""
    d = str(tmp_path)
    tabs = _kat_tables(d)
    small = bamio.encode_record(kats.se("decoy", 0, 5100, 37))
    payload = small * (payload_kb * 1024 // len(small))
    carrier = kats.se("carrier", 0, 1050, 37, aux=[("NM", "C", 0), ("ZB", "B", b"C" + struct.pack("<i", len(payload)) + payload)])
    reads = [kats.se("a%d" % j, 16 if j % 2 else 0, 1000 + 7 * j, 30 + j % 3, aux=[("NM", "C", 1)]) for j in range(300)]
    reads.append(carrier)
    reads += [kats.se("b%d" % j, 16 if j % 2 else 0, 1100 + 5 * j, 30 + j % 3, aux=[("NM", "C", 1)]) for j in range(300)]
    bam = os.path.join(d, "reads.bam")
    raw = bamio.write_bam(bam, [("chr1", 1000000)], reads, level=1)
    ora = O.OracleIndex(*tabs)
    want = ora.scan_stream(raw, O.default_opts())
    assert want[0] == len(reads) and want[9] > 0
    whole = capi.Index(*tabs)
    assert whole.scan_alignments(bam, capi.default_opts()) == want
    ixs, reps, redone = run_parts(tabs, bam, 2, capi.default_opts())
    assert redone == [1]                                       # the guess was wrong, and it was noticed
    assert reps[1][0] == reps[0][1] and reps[0][1] > 0         # the second part now starts where the carrier record ends
    assert_parts_equal_whole(ixs, whole)
    for ix in ixs + [whole]:
        ix.close()
    ora.close()


def test_a_record_longer_than_a_whole_part(tmp_path):
    """five ranks, one 3 MB record in the middle of a small file: the parts it runs over entirely hold no record start, say so
    (or are told so), and contribute nothing"""
    d = str(tmp_path)
    tabs = _kat_tables(d)
    payload = bytes(range(256)) * (3 * 4096)                    # no structure to take for a record
    reads = [kats.se("a%d" % j, 0, 1000 + 7 * j, 37, aux=[("NM", "C", 1)]) for j in range(2000)]
    reads.append(kats.se("long", 0, 1050, 37, aux=[("ZB", "B", b"C" + struct.pack("<i", len(payload)) + payload)]))
    reads += [kats.se("b%d" % j, 16, 1100 + 5 * j, 37, aux=[("NM", "C", 1)]) for j in range(2000)]
    bam = os.path.join(d, "reads.bam")
    raw = bamio.write_bam(bam, [("chr1", 1000000)], reads, level=0)      # stored blocks: file offsets follow stream offsets
    ora = O.OracleIndex(*tabs)
    want = ora.scan_stream(raw, O.default_opts())
    whole = capi.Index(*tabs)
    assert whole.scan_alignments(bam, capi.default_opts()) == want
    assert whole.profile()["n_replayed_windows"] == 0           # spans under the long record are consistent as they stand
    ixs, reps, redone = run_parts(tabs, bam, 5, capi.default_opts())
    assert any(int(ix.cnt[0]) == 0 for ix in ixs[1:-1])         # at least one part lies inside the record
    assert_parts_equal_whole(ixs, whole)
    for ix in ixs + [whole]:
        ix.close()
    ora.close()


def test_more_ranks_than_blocks_and_unsupported_modes(tmp_path):
    d = str(tmp_path)
    tabs = _kat_tables(d)
    reads = [kats.se("a%d" % j, 0, 1000 + 7 * j, 37) for j in range(50)]
    bam = os.path.join(d, "reads.bam")
    bamio.write_bam(bam, [("chr1", 1000000)], reads, level=6)          # one data block + the end-of-file block
    whole = capi.Index(*tabs)
    want = whole.scan_alignments(bam, capi.default_opts())
    ixs, reps, _ = run_parts(tabs, bam, 6, capi.default_opts())
    assert [sum(int(ix.cnt[k]) for ix in ixs) for k in range(13)] == want
    assert sum(1 for r in reps if r[2]) == 1                    # one rank owns the only block
    with pytest.raises(capi.ItxError):
        ixs[1].scan_shard_file(bam, capi.default_opts(rmDup=1), 1, 2)
    with pytest.raises(capi.ItxError):
        ixs[1].scan_shard_file(bam, capi.default_opts(), 2, 2)
    for ix in ixs + [whole]:
        ix.close()


def test_cpg_parts_sum_to_the_whole(tmp_path):
    """cpgstat over line-aligned parts of the bedGraph: counts exact, score sums within 1e-9 relative"""
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=23)
    tabs = s.write_tables(d)
    bg = os.path.join(d, "cpg.bedGraph")
    s.write_bedgraph(bg, 200000)
    whole = capi.Index(*tabs)
    lines, inrep = whole.scan_cpg(bg)
    nranks = 3
    ixs = [capi.Index(*tabs) for _ in range(nranks)]
    got = [ixs[r].scan_cpg_shard(bg, r, nranks) for r in range(nranks)]
    assert sum(g[0] for g in got) == lines and sum(g[1] for g in got) == inrep and all(g[0] > 0 for g in got)
    L = whole.L
    c4 = (C.c_uint64 * 4)()
    for ix in ixs + [whole]:
        ix.sync()
    for which in range(3):
        for i in range(whole.n(which)):
            pass
    # per-subfamily CpG counts and score sums through the writers
    for k, ix in enumerate(ixs + [whole]):
        ix.write_cpg_stat(os.path.join(d, "p%d" % k))
    def rows(prefix):
        out = {}
        for ln in open(prefix + ".CpG.subfamily.stat").read().split("\n")[1:]:
            f = ln.split("\t")
            if len(f) > 3:
                out[f[0]] = f
        return out
    w = rows(os.path.join(d, "p%d" % nranks))
    parts = [rows(os.path.join(d, "p%d" % k)) for k in range(nranks)]
    hdr = open(os.path.join(d, "p0.CpG.subfamily.stat")).readline().rstrip("\n").split("\t")
    ci = [i for i, h in enumerate(hdr) if "count" in h.lower() or "cpg" in h.lower()]
    assert w and ci
    for name, f in w.items():
        for i in range(1, len(f)):
            try:
                tot = sum(float(p[name][i]) for p in parts)
                ref = float(f[i])
            except ValueError:
                continue
            if hdr[i].lower() in ("length", "genome_count", "total_length") or all(p[name][i] == f[i] for p in parts):
                continue                                        # per-subfamily constants, not sums
            if "mean" in hdr[i].lower() or "avg" in hdr[i].lower() or "density" in hdr[i].lower() or "per" in hdr[i].lower():
                continue                                        # ratios do not add
            assert abs(tot - ref) <= 1e-9 * max(abs(ref), 1.0) + 1.6e-4 * nranks, (name, hdr[i], tot, ref)
    for ix in ixs + [whole]:
        ix.close()
    s.close()


@pytest.mark.parametrize("api", ["file", "pinned"])
def test_compressed_ring_wraps(api, tmp_path, monkeypatch):
    """a compressed ring much smaller than the file (1 MiB windows, 4 MiB ring, groups of 64 blocks): windows land on bytes whose
    inflate groups are done, open groups that are in the way are closed early -- and the stream, the counts and the tables are
    those of the unbounded run"""
    d = str(tmp_path)
    s = synth.Synth(1, 60000, seed=24)
    tabs = s.write_tables(d)
    bam = os.path.join(d, "reads.bam")
    n, nrec = s.write_bam(bam, 2, 150000, level=1, threads=4)
    assert os.path.getsize(bam) > 12 << 20
    whole = capi.Index(*tabs)
    want = whole.scan_alignments(bam, capi.default_opts())
    stream_want = whole.stream_fetch(n)
    monkeypatch.setenv("ITX_COMP_WINDOW_MB", "1")
    monkeypatch.setenv("ITX_COMP_RING_MB", "4")
    monkeypatch.setenv("ITX_INF_GROUP", "64")
    ix = capi.Index(*tabs)
    if api == "file":
        got = ix.scan_alignments(bam, capi.default_opts())
    else:
        L = capi.lib()
        fsz = os.path.getsize(bam)
        pin = L.itx_host_alloc_pinned(fsz + 64)
        with open(bam, "rb") as f:
            assert f.readinto((C.c_char * fsz).from_address(pin)) == fsz
        got = ix.scan_bgzf_memory(pin, fsz, capi.default_opts())
        L.itx_host_free_pinned(pin)
    assert got == want and got[0] + got[1] == nrec
    assert ix.stream_fetch(n) == stream_want                    # what the device inflated, byte for byte
    for which in range(3):
        assert ix.table(which) == whole.table(which)
    ix.close(); whole.close(); s.close()
