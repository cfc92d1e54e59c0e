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
