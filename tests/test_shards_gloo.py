"""Multi-rank path on the CPU: two gloo ranks each run the device logic (host emulation) on their own
record-range shard of one coordinate-sorted stream and sum-reduce the packed integer counter block;
the result must equal the single-rank scan of the whole stream, whatever the reduction order
(u64 sums; u32 coverage difference arrays wrap like the reference's unsigned int)."""
import os
import subprocess
import sys
import textwrap

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path[:0] = [%(root)r, %(here)r]
    import numpy as np, torch, torch.distributed as dist
    import emu_lib, synth
    from iteres_b200 import capi
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    d = sys.argv[1]
    s = synth.Synth(1, 30000, seed=11)
    tabs = [os.path.join(d, n) for n in ("chrom.sizes", "rep.sizes", "rmsk.txt")]
    if rank == 0:
        s.write_tables(d)
    dist.barrier()
    n_units, mode = 40000, 2
    nch = s.n_chunks(n_units)
    c0, c1 = rank * nch // world, (rank + 1) * nch // world
    buf, n, nrec = s.stream(mode, n_units, c0, c1)
    ix = emu_lib.EmuIndex(*tabs, chunk=2048)
    cnt = ix.scan_stream(buf[:n].tobytes(), capi.default_opts())
    # the packed counter block: 13 u64 + group pairs (as int64 bit patterns), coverage as u32 -> int64 lanes mod 2^32
    t = ix.table(0) + ix.table(1) + ix.table(2)
    u64 = torch.tensor(list(cnt) + [v for row in t for v in row[1:3]], dtype=torch.int64)
    cov = np.concatenate([ix.coverage(i, u) for i in range(ix.n(0)) for u in (0, 1)]).astype(np.int64)
    cov_t = torch.from_numpy(cov)
    dist.all_reduce(u64); dist.all_reduce(cov_t)
    cov_sum = (cov_t.numpy() %% (1 << 32)).astype(np.uint32)
    if rank == 0:
        whole, n2, nrec2 = s.stream(mode, n_units)
        ref = emu_lib.EmuIndex(*tabs, chunk=2048)
        c_ref = ref.scan_stream(whole[:n2].tobytes(), capi.default_opts())
        t_ref = ref.table(0) + ref.table(1) + ref.table(2)
        want = list(c_ref) + [v for row in t_ref for v in row[1:3]]
        cov_ref = np.concatenate([ref.coverage(i, u) for i in range(ref.n(0)) for u in (0, 1)])
        ok = (u64.tolist() == want) and np.array_equal(cov_sum, cov_ref) and c_ref[9] > 0
        print(json.dumps({"ok": bool(ok), "records": int(u64[0] + u64[1]), "in_repeats": int(u64[9])}))
    dist.destroy_process_group()
""")


def test_two_rank_shards_sum_to_the_whole(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, here=HERE))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script), str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["records"] == 80000, res


# ------------------------------------------------------------------------------------------ ONE stream cut across ranks at arbitrary bytes
SHARD_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path[:0] = [%(root)r, %(here)r]
    import numpy as np, torch, torch.distributed as dist
    import emu_lib, synth
    from iteres_b200 import capi
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    d = sys.argv[1]
    s = synth.Synth(1, 30000, seed=12)
    tabs = [os.path.join(d, n) for n in ("chrom.sizes", "rep.sizes", "rmsk.txt")]
    if rank == 0:
        s.write_tables(d)
    dist.barrier()
    whole, n, nrec = s.stream(2, 30000)
    raw = whole[:n].tobytes()
    hdr_len = len(s.header())
    results = []
    for cut_seed in range(4):
        # the ranks' parts: byte ranges of the record stream cut at ARBITRARY positions (the middle of a record, of its core, of a
        # pair), as BGZF block boundaries fall; every rank reads 4 KiB past its own end for the record that straddles it
        rng = np.random.RandomState(100 + cut_seed)
        cuts = [0] + sorted(int(x) for x in rng.randint(hdr_len + 1, n, size=world - 1)) + [n]
        a, b = cuts[rank], cuts[rank + 1]
        part = raw[a:min(n, b + 4096)]
        own = b - a
        ix = emu_lib.EmuIndex(*tabs, chunk=4096)
        carry0 = hdr_len if rank == 0 else capi.SHARD_GUESS
        for rounds in range(world + 2):
            ix.reset()
            cnt, en, ex = ix.scan_shard(raw[:hdr_len + 64], part, own, carry0, capi.default_opts())
            # (u64 bit patterns in int64 lanes)
            rep = torch.from_numpy(np.array([en, ex, own], dtype=np.uint64).view(np.int64).copy())
            allr = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(allr, rep)
            reps = [tuple(int(v) for v in t.numpy().view(np.uint64)) for t in allr]
            bad, forced = capi.shard_chain_check(reps, emu_lib.lib())
            if bad < 0:
                break
            if bad == rank:
                carry0 = forced
        else:
            raise SystemExit("the parts do not chain")
        t = ix.table(0) + ix.table(1) + ix.table(2)
        u64 = torch.tensor(list(cnt) + [v for row in t for v in row[1:3]], dtype=torch.int64)
        cov_t = torch.from_numpy(np.concatenate([ix.coverage(i, u) for i in range(ix.n(0)) for u in (0, 1)]).astype(np.int64))
        dist.all_reduce(u64); dist.all_reduce(cov_t)
        cov_sum = (cov_t.numpy() %% (1 << 32)).astype(np.uint32)
        if rank == 0:
            ref = emu_lib.EmuIndex(*tabs, chunk=4096)
            c_ref = ref.scan_stream(raw, capi.default_opts())
            t_ref = ref.table(0) + ref.table(1) + ref.table(2)
            want = list(c_ref) + [v for row in t_ref for v in row[1:3]]
            cov_ref = np.concatenate([ref.coverage(i, u) for i in range(ref.n(0)) for u in (0, 1)])
            results.append(bool(u64.tolist() == want and np.array_equal(cov_sum, cov_ref) and c_ref[9] > 0))
    if rank == 0:
        print(json.dumps({"ok": all(results), "cases": len(results)}))
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("world", [2, 3])
def test_one_stream_cut_mid_record_across_ranks(world, tmp_path):
    """the protocol of itx_scan_alignments_shard on the CPU: every rank scans its byte range of ONE stream from a GUESSED first record
    (the host emulation of the device logic), the reports are all-gathered over gloo, itx_shard_chain_check decides who scans again,
    and the summed counter blocks are those of a single pass"""
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER % dict(root=ROOT, here=HERE))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
                        "--master-port", str(29540 + world), str(script), str(tmp_path)], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["cases"] == 4, res
