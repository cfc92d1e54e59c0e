#!/usr/bin/env python
"""bench.py -- the iteres hot path on B200: aligned reads/sec for `iteres stat` (rmsk overlap count).

One step = one complete pass of the hot path over the workload: reset counters -> record-boundary +
decode kernel -> overlap/selection/accumulation kernel (-> one NCCL allreduce of the counter block
when N > 1) -> the 13 global counters back on the host.

  value      reads/s with the uncompressed BAM stream already resident in HBM (kernels only)
  e2e        the same metric through the public C-ABI call a user makes, itx_scan_alignments() on a
             BGZF .bam file: host threads pread() the compressed file into pinned memory window by window
             -> cudaMemcpyAsync -> k_inflate + k_lz_resolve (BGZF blocks inflated on the device in groups
             pipelined over several streams; ITX_INFLATE=host keeps zlib on the host threads) -> scan
             kernels -> counter tables back on the host, all inside the timed region
  roofline   the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref/iteres stat) on the box's host cores,
             on a bounded sample of the same workload (reference is single threaded by construction;
             P independent copies are run side by side, P = cores used)

N = 1 workload: BASELINE.json configs[1] (50 M SE-50 reads, hg19-shaped, vs a 5.5 M row rmsk table).
N > 1: weak scaling, every rank scans its own 50 M read genomic-coordinate shard of an N x 50 M read
coordinate-sorted stream (BASELINE.json configs[4]'s sharding), merged by one allreduce per step.
Data is synthetic (tools/itx_synth.c), generated on the box; nothing is read from /root/reference.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "aligned reads/sec for iteres stat (rmsk overlap count); HBM GB/s vs peak"
UNIT = "reads/s"


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.rows, self.gpu, self.p = [], gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.perf_counter()])

    def wait_first(self, timeout=4.0):
        t0 = time.perf_counter()
        while self.p and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                pass
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", 1e30)
        inside = [r for r in self.rows if t0 - 0.15 <= r[-1] <= t1 + 0.15]
        self.rows = inside or self.rows
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def workdir():
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(cand).free > (24 << 30):
                base = cand
                break
        except Exception:
            pass
    base = base or tempfile.gettempdir()
    d = os.path.join(base, "itx_bench_%s" % os.environ.get("MASTER_PORT", str(os.getpid())))
    os.makedirs(d, exist_ok=True)
    return d


# ------------------------------------------------------------------------------------------ reference arm
def reference_baseline(wd, tables, synth_world, sample_reads, mode, procs, repeats=1):
    """Time the unmodified reference (oracle/_ref/iteres stat) on `procs` host cores: each process scans
    its own sample BAM of sample_reads reads (same generator, same rmsk).  The fixed cost (rmsk parse,
    wig/bigWig writing) is measured with header-only BAMs and subtracted, so the figure is the read loop."""
    import oracle_lib as O
    cs, rs, rm = tables
    kind = "reference" if os.path.exists(O.REF_BIN) else "port"
    hdr = synth_world.header()
    empty = os.path.join(wd, "empty.bam")
    import synth as S
    S.lib().synth_write_bam(empty.encode(), hdr.ctypes.data, len(hdr), hdr.ctypes.data, 0, 1, 1)
    bams = []
    for i in range(procs):
        b = os.path.join(wd, "sample%d.bam" % i)
        if not os.path.exists(b):
            keep = synth_world.seed
            synth_world.seed = 1000 + i          # same tables (built from the construction seed), different reads per process
            synth_world.write_bam(b, mode, sample_reads, level=1, threads=max(1, (os.cpu_count() or 8) // 2))
            synth_world.seed = keep
        bams.append(b)

    def run_all(paths, tag):
        t0 = time.perf_counter()
        ps = []
        for i, b in enumerate(paths):
            od = os.path.join(wd, "ref_%s_%d" % (tag, i))
            os.makedirs(od, exist_ok=True)
            if kind == "reference":
                ps.append(subprocess.Popen([O.REF_BIN, "stat", "-o", "out", cs, rs, rm, b], cwd=od, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            else:
                code = ("import sys; sys.path[:0]=[%r,%r]; import oracle_lib as O; ix=O.OracleIndex(%r,%r,%r); ix.scan_file(%r,O.default_opts()); ix.write_stat('out')"
                        % (ROOT, os.path.join(ROOT, "tests"), cs, rs, rm, b))
                ps.append(subprocess.Popen([sys.executable, "-c", code], cwd=od, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        rcs = [p.wait() for p in ps]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference run failed: %r" % rcs)
        return dt

    fixed = run_all([empty] * procs, "fixed")
    times = [run_all(bams, "full") for _ in range(repeats)]
    loop = [max(t - fixed, 1e-6) for t in times]
    return {"kind": kind, "cores": procs, "sample_reads": sample_reads * procs, "fixed_s": fixed, "wall_s": times, "loop_s": loop,
            "sample": "%d x %d reads (same generator and rmsk table as the GPU arm, %s), one single-threaded `iteres stat` per core; "
                      "fixed cost (rmsk parse + wig/bigWig, header-only BAM) of %.1f s subtracted" % (procs, sample_reads, "SE-50 hg19-shaped", fixed)}


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line of the contract, on the process's original stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    # libraries (NCCL prints its version banner with some NCCL_DEBUG settings) must not add lines to stdout: everything
    # written to fd 1 during the run goes to stderr; the JSON line is written to a duplicate of the original stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads per GPU (BASELINE configs[1])")
    ap.add_argument("--rmsk", type=int, default=5_500_000)
    ap.add_argument("--mode", type=int, default=0, help="0 SE-50 (configs[1]), 1 SE-75 + XA, 2 PE-100")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="reads per reference process in the CPU baseline")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if a.warmup < 3 and a.impl == "ours":
        log("note: fewer than 3 warm-up steps requested")
    ncpu = os.cpu_count() or 8
    import synth as S
    wd = workdir()
    shape = 1
    workload = "iteres stat, %d M %s reads per GPU, hg19-shaped, vs %.1f M-interval synthetic rmsk" % (
        a.reads // 1_000_000, {0: "SE-50", 1: "SE-75+XA", 2: "PE-100"}[a.mode], a.rmsk / 1e6)

    # ---------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        world_s = S.Synth(shape, a.rmsk, seed=1)
        tables = world_s.write_tables(os.path.join(wd, "tables"))
        procs = a.cpu_procs or min(ncpu, 8)
        vals = []
        res = None
        for i in range(a.warmup + a.steps):
            res = reference_baseline(wd, tables, world_s, a.cpu_sample, a.mode, procs)
            if i >= a.warmup:
                vals.append(res["sample_reads"] / res["loop_s"][0])
        v = sum(vals) / len(vals)
        ms = 1e3 * res["sample_reads"] / v
        emit(({"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer (f32 coverage ratio)",
                          "data": "synthetic", "config": {"workload": workload, "step": "bounded sample: " + res["sample"]},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": res["kind"], "sample": res["sample"]},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        if not a.keep:
            shutil.rmtree(wd, ignore_errors=True)
        return 0

    # ---------------------------------------------------------------- our arm
    import iteres_b200 as itx
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    L = itx.lib()
    if L.itx_device_count() <= lrank:
        raise SystemExit("bench.py: no CUDA device %d (the product has no CPU path)" % lrank)

    def barrier():
        if dist:
            dist.barrier()

    t_setup = time.perf_counter()
    world_s = S.Synth(shape, a.rmsk, seed=1)
    tdir = os.path.join(wd, "tables")
    if rank == 0:
        tables = world_s.write_tables(tdir)
    barrier()
    tables = tuple(os.path.join(tdir, n) for n in ("chrom.sizes", "rep.sizes", "rmsk.txt"))
    ix = itx.Index(*tables, device=lrank)
    if a.chunk or a.window:
        ix.tune(chunk_bytes=a.chunk, window_bytes=a.window)
    if world > 1:
        ix.tune(inflate_threads=max(2, (os.cpu_count() or 16) // world))      # the ranks of a box share its host cores
    log("index: %d intervals, %d subfamilies (%.1f s)" % (L.itx_n_elem(ix.h), ix.n(0), time.perf_counter() - t_setup))
    if world > 1:
        uid = [ix.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ix.comm_init(uid[0], rank, world)

    # this rank's shard of the coordinate-sorted stream: generator chunks [c0, c1)
    n_units = a.reads * world
    nch = world_s.n_chunks(n_units)
    c0, c1 = rank * nch // world, (rank + 1) * nch // world
    gth = max(1, ncpu // world)
    hdr = world_s.header()
    sz, nrec = world_s.records_size(a.mode, n_units, c0, c1, gth)
    n = len(hdr) + sz
    hbuf = L.itx_host_alloc_pinned(n + 64)
    if not hbuf:
        raise SystemExit("pinned allocation of %d bytes failed" % (n + 64))
    C.memmove(hbuf, hdr.ctypes.data, len(hdr))
    got = world_s.records_into(hbuf + len(hdr), a.mode, n_units, c0, c1, gth)
    assert got == sz
    C.memset(hbuf + n, 0, 64)
    dbuf = L.itx_dev_alloc(n + 64)
    assert dbuf and L.itx_dev_upload(dbuf, hbuf, n + 64) == 0
    h = ix.header(hbuf, n)
    opts = itx.default_opts()
    log("shard: %d records, %.2f GB uncompressed, generated+uploaded (%.1f s since start)" % (nrec, n / 1e9, time.perf_counter() - t_setup))

    def step():
        ix.reset()
        cnt = ix.scan_bam_device(h, dbuf, n, opts)     # this rank's counters
        if world > 1:
            step.global_cnt = ix.allreduce_counts()    # the job's counters after the one allreduce
        return cnt
    step.global_cnt = None

    for _ in range(a.warmup):
        step()
    L.itx_dev_sync()
    barrier()
    sampler = ClockSampler(lrank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    barrier()
    for _ in range(2):              # every rank: keep the devices busy until the sampler is running (untimed)
        step()
    L.itx_dev_sync()
    barrier()
    dec = ovl = 0.0
    launches = 0
    t_w0 = time.perf_counter()
    ix.mark(0)
    for _ in range(a.steps):
        cnt = step()
        pr = ix.profile()
        dec += pr["decode_ms"]; ovl += pr["overlap_ms"]; launches += pr["n_launches"]
    ix.mark(1)
    L.itx_dev_sync()
    el_ms = ix.elapsed_ms(0, 1)
    sampler.window(t_w0, time.perf_counter())
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    assert cnt[0] + cnt[1] == nrec, (cnt, nrec)
    if dist:
        import torch
        t = torch.tensor([el_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el_ms = float(t[0])
        tot = torch.tensor([float(nrec)], dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_rec = int(tot[0])
        assert step.global_cnt[0] + step.global_cnt[1] == total_rec, (step.global_cnt, total_rec)   # the allreduce summed every shard
    else:
        total_rec = nrec
    value = total_rec * a.steps / (el_ms * 1e-3)
    pr = ix.profile()
    bad = pr["n_bad_chunks"]

    # roofline of the dominant kernel (per step, this rank)
    peak, peak_src = peaks()
    R, F, H, HU = cnt[0] + cnt[1], cnt[6], cnt[9] + cnt[12], cnt[10]
    k1_bytes = n + 16 * R
    k2_bytes = 16 * R + 16 * F + 16 * H + (4 + 8) * H + 8 * HU
    dec_ms, ovl_ms = dec / a.steps, ovl / a.steps
    fused = bool(pr.get("fused"))
    all_k = {}
    if fused:
        # one kernel per launch group: the stream is read once, no tuple leaves the SM
        ks_bytes = n + 16 * F + 16 * H + (4 + 8) * H + 8 * HU
        dom = ("k_scan (K1+K2+K3 fused: record boundaries, decode, overlap, selection, accumulation)", ks_bytes, dec_ms)
        all_k["k_scan"] = {"ms": dec_ms, "bytes": ks_bytes, "GBps": ks_bytes / max(dec_ms, 1e-9) / 1e6, "frac": ks_bytes / max(dec_ms, 1e-9) / 1e6 / peak}
        # the tuple path (what -R and the ordered outputs run), timed outside the timed region for comparison
        os.environ["ITX_FUSED"] = "0"
        d2 = o2 = 0.0
        for _ in range(5):
            step()
            p2 = ix.profile()
            d2 += p2["decode_ms"] / 5; o2 += p2["overlap_ms"] / 5
        del os.environ["ITX_FUSED"]
        all_k["tuple_path"] = {"k_decode_span": {"ms": d2, "bytes": k1_bytes, "GBps": k1_bytes / max(d2, 1e-9) / 1e6, "frac": k1_bytes / max(d2, 1e-9) / 1e6 / peak},
                               "k_overlap": {"ms": o2, "bytes": k2_bytes, "GBps": k2_bytes / max(o2, 1e-9) / 1e6, "frac": k2_bytes / max(o2, 1e-9) / 1e6 / peak}}
    else:
        dom = ("k_decode_span (K1: record boundaries + decode, incl. chain verify)", k1_bytes, dec_ms) if dec_ms >= ovl_ms else \
              ("k_overlap (K2+K3: interval overlap, selection, accumulation)", k2_bytes, ovl_ms)
        all_k = {"k_decode_span": {"ms": dec_ms, "bytes": k1_bytes, "GBps": k1_bytes / max(dec_ms, 1e-9) / 1e6, "frac": k1_bytes / max(dec_ms, 1e-9) / 1e6 / peak},
                 "k_overlap": {"ms": ovl_ms, "bytes": k2_bytes, "GBps": k2_bytes / max(ovl_ms, 1e-9) / 1e6, "frac": k2_bytes / max(ovl_ms, 1e-9) / 1e6 / peak}}
    ach = dom[1] / (dom[2] * 1e-3) / 1e9 if dom[2] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("reads_per_gpu") == a.reads and tj.get("mode") == a.mode:
            traffic = tj.get(dom[0].split(" ")[0])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch_group": dom[1], "kernel_ms_per_step": dom[2],
                "replayed_windows": int(pr.get("n_replayed_windows", 0)), "all_kernels": all_k}

    # ---------------------------------------------------------------- e2e: BGZF file -> tables, through itx_scan_alignments
    e2e = None
    if not a.no_e2e:
        bam = os.path.join(wd, "shard%d.bam" % rank)
        t0 = time.perf_counter()
        rc = S.lib().synth_write_bam(bam.encode(), hbuf, len(hdr), hbuf + len(hdr), sz, 1, gth)
        assert rc == 0
        log("BGZF shard written: %.2f GB (%.1f s)" % (os.path.getsize(bam) / 1e9, time.perf_counter() - t0))
        if a.chunk or a.window:
            pass
        times = []
        for i in range(1 + a.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            ix.reset()
            t_a = time.perf_counter()
            c2 = ix.scan_alignments(bam, opts)
            t_b = time.perf_counter()
            if world > 1:
                ix.allreduce_counts()
            ix.sync()                                   # counter tables + coverage vectors back on the host
            dt = time.perf_counter() - t0
            log("e2e step %d: reset %.1f ms, scan %.1f ms, allreduce+sync %.1f ms" % (i, 1e3 * (t_a - t0), 1e3 * (t_b - t_a), 1e3 * (t0 + dt - t_b)))
            if dist:
                import torch
                tt = torch.tensor([dt], dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt[0])
            if i:
                times.append(dt)
            if world == 1:
                assert c2 == cnt, (c2, cnt)
        pe = ix.profile()
        e2e_t = sum(times) / len(times)
        e2e = {"value": total_rec / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(pe["h2d_bytes"]), "d2h_bytes_per_step": int(pe["d2h_bytes"]),
               "s_per_step": e2e_t, "steps": len(times), "api": "itx_scan_alignments(BGZF .bam, deflate level 1) + itx_sync_counts",
               "inflate": ("host zlib threads" if pe["inflate_threads"] else "device: k_inflate (Huffman pass, one thread per BGZF block) + k_lz_resolve (match copies, one CTA per block), groups of 16384 blocks on 8 streams"),
               "inflate_threads": int(pe["inflate_threads"]), "inflate_ms": pe["inflate_ms"],
               "scan_stream_ms": pe["decode_ms"] + pe["overlap_ms"],      # event time on the scan stream: includes its waits on the inflate streams
               "bam_bytes": os.path.getsize(bam)}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            procs = a.cpu_procs or min(ncpu, 8)
            r = reference_baseline(wd, tables, world_s, a.cpu_sample, a.mode, procs)
            cpu = {"value": r["sample_reads"] / r["loop_s"][0], "unit": UNIT, "cores": procs, "kind": r["kind"], "sample": r["sample"],
                   "wall_s": r["wall_s"][0], "fixed_s": r["fixed_s"], "host_cpus": ncpu}
        except Exception as ex:   # the baseline is reported, never load-bearing
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "failed: %s" % ex}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": el_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u32/u64 integer (f32 coverage ratio)", "data": "synthetic",
               "config": {"workload": workload, "records_per_gpu": nrec, "stream_bytes_per_gpu": n, "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2; no flush needed" % (n / 1e9),
                          "chunk_bytes": a.chunk or 65536, "parallelism": "genome-coordinate shards x%d, one allreduce of the counter block per step" % world if world > 1 else "single GPU",
                          "repaired_chunk_entries": int(bad)},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "counters": {"records": int(cnt[0] + cnt[1]), "fragments": int(cnt[6]), "in_repeats": int(cnt[9]), "unique_in_repeats": int(cnt[10])}}
        emit(out)
    L.itx_bam_header_free(h)
    L.itx_dev_free(dbuf)
    L.itx_host_free_pinned(hbuf)
    ix.close()
    barrier()
    if rank == 0 and not a.keep:
        shutil.rmtree(wd, ignore_errors=True)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
