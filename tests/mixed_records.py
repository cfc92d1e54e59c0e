"""A stream whose records range from 60 bytes to 70 KB (long reads, long qnames, many CIGAR ops, big aux
arrays): records larger than a decode chunk, than the kernel's shared-memory ring and than a copy window,
records that end exactly on tile boundaries, and XA tags behind big aux fields."""
import random

import bamio


def make(n=1500, seed=3, refs=(("chr1", 249250621), ("chr2", 243199373), ("chrUn_gl000220", 161802))):
    rnd = random.Random(seed)
    recs = []
    lens = [20, 36, 50, 100, 150, 151, 250, 600, 1500, 1900, 2040, 3000, 4090, 9000, 20000, 70000]
    for i in range(n):
        L = rnd.choice(lens) if rnd.random() < 0.35 else rnd.choice(lens[:6])
        tid = rnd.choice([0, 0, 0, 1, 1, 2, -1])
        flag = rnd.choice([0, 16, 0, 16, 4, 99, 147, 83, 163, 73, 133])
        if tid < 0:
            flag |= 4
        pos = rnd.randrange(0, 3_000_000) if tid >= 0 else -1
        nops = rnd.choice([1, 1, 1, 3, 5, 40])
        if flag & 4:
            cigar = "*"
        elif nops == 1:
            cigar = "%dM" % L
        else:
            per = max(1, L // nops)
            cigar = "".join("%d%s" % (per, "MDMNMSMIM=X"[k % 11]) for k in range(nops))
        aux = [("NM", "C", rnd.randrange(4))]
        if rnd.random() < 0.2:
            aux.append(("ZB", "B", b"S" + (300).to_bytes(4, "little") + bytes(600)))
        if rnd.random() < 0.3:
            aux += [("NM", "i", rnd.randrange(3)), ("XA", "Z", "chr1,%s%d,%dM,%d;chr2,-%d,36M,0;" % (
                rnd.choice("+-"), rnd.randrange(1, 3_000_000), L, rnd.randrange(3), rnd.randrange(1, 3_000_000)))]
        if rnd.random() < 0.1:
            aux.insert(0, ("XF", "f", 1.5))
        q = "r%d" % i + ("_long_name" * rnd.choice([0, 0, 0, 20]))
        recs.append(dict(qname=q[:250], flag=flag, tid=tid, pos=pos, mapq=rnd.choice([0, 3, 20, 37, 60]), cigar=cigar,
                         mtid=tid, mpos=max(0, pos + rnd.randrange(-400, 400)) if tid >= 0 else -1,
                         isize=rnd.choice([0, 200, -200, 480, -480, 700]), seq="A" * L, qual="I" * L, aux=aux))
    return bamio.encode_header(list(refs)) + b"".join(bamio.encode_record(r) for r in recs), len(recs)


def tables(d, n_el=4000, seed=5):
    """chrom sizes / repeat sizes / rmsk covering the first 3 Mb of chr1 and chr2 densely"""
    import os
    rnd = random.Random(seed)
    with open(os.path.join(d, "chrom.sizes"), "w") as f:
        f.write("chr1\t249250621\nchr2\t243199373\n")
    names = ["SUB%03d" % i for i in range(40)]
    with open(os.path.join(d, "rep.sizes"), "w") as f:
        f.write("".join("%s\t%d\n" % (nm, 300 + 50 * i) for i, nm in enumerate(names) if i % 7))
    rows = []
    for c in ("chr1", "chr2"):
        p = 0
        for _ in range(n_el):
            p += rnd.randrange(50, 1200)
            ln = rnd.choice([30, 120, 300, 900, 5000])
            s = rnd.randrange(len(names))
            strand = rnd.choice("+-")
            cs = rnd.randrange(0, 200)
            cols = (cs, cs + ln, -5) if strand == "+" else (-5, cs + ln, cs)
            rows.append("585\t1000\t10\t5\t5\t%s\t%d\t%d\t-1\t%s\t%s\tCLS%d\tFAM%d\t%d\t%d\t%d\t1" % (
                c, p, p + ln, strand, names[s], s % 4, s % 9, cols[0], cols[1], cols[2]))
            if rnd.random() < 0.1:
                rows.append("585\t1000\t10\t5\t5\t%s\t%d\t%d\t-1\t+\t%s\tCLS%d\tFAM%d\t1\t%d\t0\t1" % (
                    c, p + 10, p + 10 + ln // 2, names[(s + 1) % len(names)], (s + 1) % 4, (s + 1) % 9, ln // 2))
    with open(os.path.join(d, "rmsk.txt"), "w") as f:
        f.write("\n".join(rows) + "\n")
    return tuple(os.path.join(d, n) for n in ("chrom.sizes", "rep.sizes", "rmsk.txt"))
