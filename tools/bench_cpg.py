#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[3]): `iteres cpgstat` on a synthetic 28 M-row CpG bedGraph vs the 5.5 M
row rmsk table.  Times itx_scan_cpg through the C ABI with the text parsed on the device (k_bedgraph) and, for
comparison, with ITX_CPG_PARSE=host, and the reference binary's cpgstat on a bounded sample.  Prints one JSON line.
usage: python tools/bench_cpg.py [--rows 28000000] [--rmsk 5500000]"""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402
import iteres_b200 as itx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=28_000_000)
    ap.add_argument("--rmsk", type=int, default=5_500_000)
    ap.add_argument("--ref-rows", type=int, default=2_000_000)
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="itx_cpg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    s = synth.Synth(1, a.rmsk, seed=1)
    tables = s.write_tables(d)
    bg = os.path.join(d, "cpg.bedGraph")
    s.write_bedgraph(bg, a.rows)
    small = os.path.join(d, "cpg_small.bedGraph")
    s.write_bedgraph(small, a.ref_rows)
    ix = itx.Index(*tables, device=0)
    out = {"metric": "CpG bedGraph rows/s for iteres cpgstat (text -> tables)", "rows": a.rows, "rmsk_rows": a.rmsk, "bedgraph_bytes": os.path.getsize(bg)}
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["ITX_CPG_PARSE"] = "host"
        ts = []
        for i in range(3):
            ix.reset()
            t0 = time.perf_counter()
            n = ix.scan_cpg(bg, 0)
            ix.sync()
            ts.append(time.perf_counter() - t0)
        os.environ.pop("ITX_CPG_PARSE", None)
        out["parse_on_" + mode] = {"s": min(ts[1:]), "rows_per_s": a.rows / min(ts[1:]), "lines": n[0], "in_repeats": n[1]}
    ref = os.path.join(ROOT, "oracle", "_ref", "iteres")
    if os.path.exists(ref):
        empty = os.path.join(d, "one.bedGraph")
        open(empty, "w").write(open(small).readline())
        def run(f):
            t0 = time.perf_counter()
            subprocess.run([ref, "cpgstat", "-o", os.path.join(d, "r"), *tables, f], cwd=d, capture_output=True)
            return time.perf_counter() - t0
        fixed = run(empty); full = run(small)
        out["reference_1_core"] = {"rows_per_s": a.ref_rows / max(full - fixed, 1e-9), "sample_rows": a.ref_rows, "fixed_s": fixed, "wall_s": full}
    print(json.dumps(out))
    ix.close(); s.close()
    import shutil; shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
