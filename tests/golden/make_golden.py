#!/usr/bin/env python
"""Regenerate tests/golden/ from the UNMODIFIED reference binary (oracle/_ref/iteres, built by
`make -C oracle ref` from /root/reference).  Run in the build container only; the committed
outputs are what lets the same parity checks run where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bamio  # noqa: E402
import kats  # noqa: E402

REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "iteres")


def write_inputs(d, k):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "chrom.sizes"), "w") as f:
        f.write("".join("%s\t%d\n" % x for x in k["chrom"]))
    with open(os.path.join(d, "rep.sizes"), "w") as f:
        f.write("".join("%s\t%d\n" % x for x in k["rep"]))
    with open(os.path.join(d, "rmsk.txt"), "w") as f:
        f.write("".join(r + "\n" for r in k["rmsk"]))
    if k["reads"]:
        bamio.write_bam(os.path.join(d, "reads.bam"), k["refs"], k["reads"], block=600)
        with open(os.path.join(d, "reads.sam"), "w") as f:
            f.write("".join("@SQ\tSN:%s\tLN:%d\n" % x for x in k["refs"]))
            f.write("".join(bamio.sam_line(r, k["refs"]) + "\n" for r in k["reads"]))
    if "bedgraph" in k:
        with open(os.path.join(d, "cpg.bedGraph"), "w") as f:
            f.write("".join(l + "\n" for l in k["bedgraph"]))


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    for name, k in kats.KATS.items():
        d = os.path.join(HERE, name)
        shutil.rmtree(d, ignore_errors=True)
        inp = os.path.join(d, "input")
        write_inputs(inp, k)
        for vname, (cmd, args) in k["variants"].items():
            vd = os.path.join(d, vname)
            os.makedirs(vd)
            data = "cpg.bedGraph" if cmd.startswith("cpg") else ("reads.sam" if "-S" in args else "reads.bam")
            argv = [REF, cmd] + args + ["-o", "out"] + [os.path.join("..", "input", x) for x in
                                                         ("chrom.sizes", "rep.sizes", "rmsk.txt", data)]
            p = subprocess.run(argv, cwd=vd, capture_output=True, text=True)
            with open(os.path.join(vd, "cmdline.txt"), "w") as f:
                f.write("iteres %s %s\nexit=%d\n" % (cmd, " ".join(args), p.returncode))
            with open(os.path.join(vd, "stderr.txt"), "w") as f:
                f.write(p.stderr)
            print(name, vname, "exit", p.returncode, sorted(os.listdir(vd)))


if __name__ == "__main__":
    main()
