#!/usr/bin/env python
"""Wall time of the command line a user types: `iteres stat -o out chrom.sizes rep.sizes rmsk.txt reads.bam`, ours
(iteres_b200/csrc/iteres, one process, one GPU) and the reference binary (oracle/_ref/iteres, single threaded by
construction) on the same files, outputs compared byte for byte.  Prints one JSON line.
usage: python tools/bench_cli.py [--reads 50000000] [--ref-reads 5000000] [--rmsk 5500000]"""
import argparse, filecmp, json, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402


def run(exe, args, cwd):
    t0 = time.perf_counter()
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True)
    return time.perf_counter() - t0, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=50_000_000)
    ap.add_argument("--ref-reads", type=int, default=5_000_000)
    ap.add_argument("--rmsk", type=int, default=5_500_000)
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="itx_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    s = synth.Synth(1, a.rmsk, seed=1)
    tables = list(s.write_tables(d))
    big, small = os.path.join(d, "big.bam"), os.path.join(d, "small.bam")
    s.write_bam(big, 0, a.reads, level=1, threads=os.cpu_count() or 8)
    s.write_bam(small, 0, a.ref_reads, level=1, threads=os.cpu_count() or 8)
    ours, ref = os.path.join(ROOT, "iteres_b200", "csrc", "iteres"), os.path.join(ROOT, "oracle", "_ref", "iteres")
    out = {"metric": "wall seconds of `iteres stat` (rmsk parse + scan + tables + wig + bigWig + report)", "rmsk_rows": a.rmsk}
    run(ours, ["stat", "-o", "warm"] + tables + [small], d)                       # CUDA context, page cache
    t, p = run(ours, ["stat", "-o", "o_small"] + tables + [small], d)
    out["ours_small"] = {"reads": a.ref_reads, "wall_s": t, "rc": p.returncode}
    t, p = run(ours, ["stat", "-o", "o_big"] + tables + [big], d)
    out["ours_big"] = {"reads": a.reads, "wall_s": t, "rc": p.returncode, "bam_bytes": os.path.getsize(big),
                       "timeline": [l.replace("[itx timing] ", "") for l in p.stderr.splitlines() if l.startswith("[itx timing]") and "inflate group" not in l]}    # with ITX_TIMING=1
    if os.path.exists(ref):
        t, p = run(ref, ["stat", "-o", "r_small"] + tables + [small], d)
        out["reference_small"] = {"reads": a.ref_reads, "wall_s": t, "rc": p.returncode}
        same = all(filecmp.cmp(os.path.join(d, "o_small" + e), os.path.join(d, "r_small" + e), shallow=False)
                   for e in (".iteres.subfamily.stat", ".iteres.family.stat", ".iteres.class.stat", ".iteres.report", ".iteres.bigWig", ".iteres.unique.bigWig"))
        out["outputs_identical"] = bool(same)
        out["speedup_small"] = out["reference_small"]["wall_s"] / out["ours_small"]["wall_s"]
    print(json.dumps(out))
    s.close()
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
