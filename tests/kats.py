"""Known-answer inputs for the iteres hot path (SURVEY.md 4.2 KAT-1..5 plus edge cases).

The EXPECTED outputs are not written here: tests/golden/make_golden.py runs the unmodified
reference binary (oracle/_ref/iteres) on these inputs and commits what it printed under
tests/golden/<kat>/<variant>/.  Positions below are 0-based (BAM); SAM lines are derived."""

SEQ36, QUAL36 = "A" * 36, "I" * 36


def rmsk_row(chrom, start, end, strand, name, cls, fam, rep_start, rep_end, rep_left):
    return "\t".join(map(str, [585, 1000, 10, 5, 5, chrom, start, end, -1000, strand, name, cls, fam,
                               rep_start, rep_end, rep_left, 1]))


def se(qname, flag, pos, mapq, cigar="36M", tid=0, aux=(), seq=SEQ36, qual=QUAL36):
    if flag & 4:
        return dict(qname=qname, flag=flag, tid=-1, pos=-1, mapq=0, cigar="*", seq=seq, qual=qual, aux=list(aux))
    return dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cigar, seq=seq, qual=qual, aux=list(aux))


def pe(qname, flag, pos, mapq, mpos, isize, cigar="36M", tid=0, mtid=0, aux=()):
    return dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cigar, mtid=mtid, mpos=mpos, isize=isize,
                seq=SEQ36, qual=QUAL36, aux=list(aux))


ANNOT1 = [rmsk_row("chr1", 1000, 1300, "+", "AluY", "SINE", "Alu", 1, 300, 0),
          rmsk_row("chr1", 5000, 6000, "-", "L1PA2", "LINE", "L1", -100, 5900, 4900),
          rmsk_row("chr1", 5900, 6100, "+", "MIR", "SINE", "MIR", 10, 210, -50)]
REP1 = [("AluY", 300), ("L1PA2", 6000)]
CHR1M = [("chr1", 1000000)]

KATS = {}

KATS["kat1_basic"] = dict(
    chrom=CHR1M, rep=REP1, rmsk=ANNOT1, refs=CHR1M,
    reads=[se("r1", 0, 1050, 37), se("r2", 16, 5100, 5), se("r3", 0, 5880, 30), se("r4", 4, 0, 0), se("r5", 0, 989, 30)],
    variants={"default": ("stat", ["-w"]), "S": ("stat", ["-w", "-S"]), "filter_S": ("filter", ["-S", "-r"]),
              "E0": ("stat", ["-w", "-E", "0"]), "Q31": ("stat", ["-w", "-Q", "31"]),
              "N2U1": ("stat", ["-w", "-N", "2", "-U", "1"]), "N3U2": ("stat", ["-w", "-N", "3", "-U", "2"]),
              "N1": ("stat", ["-w", "-N", "1"]),
              "filter": ("filter", []), "filter_r": ("filter", ["-r"]), "filter_nAluY": ("filter", ["-n", "AluY", "-r"]),
              "filter_cLINE": ("filter", ["-c", "LINE"]), "filter_fMIR_t1": ("filter", ["-f", "MIR", "-t", "1"])})

KATS["kat2_last_ascent"] = dict(
    chrom=CHR1M, rep=[("A", 500), ("B", 500), ("C", 500)],
    rmsk=[rmsk_row("chr1", 900, 1060, "+", "A", "SINE", "Alu", 1, 161, 0),
          rmsk_row("chr1", 1070, 1090, "+", "B", "LINE", "L1", 1, 21, 0),
          rmsk_row("chr1", 1100, 1300, "+", "C", "DNA", "hAT", 1, 201, 0)],
    refs=CHR1M, reads=[se("q1", 0, 1000, 40)],
    variants={"default": ("stat", ["-w"]), "c035": ("stat", ["-w", "-c", "0.35"]), "filter_r": ("filter", ["-r"])})

KATS["kat3_hit_order"] = dict(
    chrom=CHR1M, rep=[("X", 500), ("Y", 500), ("Z", 500)],
    rmsk=[rmsk_row("chr1", 131010, 131040, "+", "Z", "SINE", "Alu", 1, 31, 0),
          rmsk_row("chr1", 131050, 131100, "+", "X", "LINE", "L1", 1, 51, 0),
          rmsk_row("chr1", 131120, 131140, "+", "Y", "DNA", "hAT", 1, 21, 0)],
    refs=CHR1M, reads=[se("q1", 0, 131000, 40)],
    variants={"default": ("stat", ["-w"]), "filter_r": ("filter", ["-r"])})

XA1 = [("NM", "i", 1), ("XA", "Z", "chr1,+5101,36M,1;")]
XA2 = [("NM", "i", 1), ("XA", "Z", "chr1,+5101,36M,2;")]
KATS["kat4_paired_xa"] = dict(
    chrom=CHR1M, rep=REP1, rmsk=ANNOT1, refs=CHR1M,
    reads=[pe("p1", 99, 1050, 30, 1214, 200), pe("p1", 147, 1214, 30, 1050, -200),
           pe("p2", 83, 5200, 5, 5000, -236), pe("p2", 163, 5000, 5, 5200, 236),
           pe("p3", 73, 1100, 30, 1100, 0), pe("p4", 99, 1050, 30, 1714, 700),
           se("r6", 0, 1050, 37, aux=XA1), se("r7", 0, 1050, 37, aux=XA2), se("r8", 16, 1250, 37, cigar="10M5D26M")],
    variants={"default": ("stat", ["-w"]), "S_B": ("stat", ["-w", "-S", "-B", "-V"]), "B": ("stat", ["-w", "-B", "-V"]), "T": ("stat", ["-w", "-T"]),
              "x": ("stat", ["-w", "-x"]), "D": ("stat", ["-w", "-D"]), "E0": ("stat", ["-w", "-E", "0"]),
              "I150": ("stat", ["-w", "-I", "150"]), "R": ("stat", ["-w", "-R"]),
              "filter_r": ("filter", ["-r"]), "filter_t2": ("filter", ["-t", "2"]), "filter_T": ("filter", ["-T", "-r"])})

KATS["kat5_cpg"] = dict(
    chrom=CHR1M, rep=REP1, rmsk=ANNOT1, refs=CHR1M, reads=[],
    bedgraph=["chr1\t1100\t1102\t2.5", "chr1\t5950\t5952\t1.25", "chr1\t998\t1000\t3", "chr1\t999\t1001\t7", "chr2\t5\t7\t1",
              "# comment", "", "chr1 1298 1300 0.125 extra"],
    variants={"cpgstat": ("cpgstat", ["-w"]), "cpgfilter": ("cpgfilter", []), "cpgfilter_nAluY": ("cpgfilter", ["-n", "AluY"]),
              "cpgfilter_t3": ("cpgfilter", ["-t", "3"])})

# Edge cases: chromosome-end clamp, absent contig, size-2 chromosome (treated as absent: cend == 1),
# nested stacks with > 8 hits, coarse-level bins (levels 1-3), zero-length rmsk row, duplicate reads,
# row whose family/class differ from the subfamily's first row, '=' / 'X' / 'N' CIGAR ops.
_stack = [rmsk_row("chr1", 20000 + 3 * i, 20400 - 3 * i, "+" if i % 2 else "-", "STK%d" % (i % 3), "DNA", "hAT",
                   (5 + i) if i % 2 else -7, 395 + i - 6 * i, (5 + i) if not i % 2 else -7) for i in range(12)]
KATS["kat6_edges"] = dict(
    chrom=[("chr1", 1000000), ("chr2", 2), ("chr3", 5000000), ("chrDup", 10), ("chrDup", 300000)],
    rep=[("AluY", 300), ("L1PA2", 6000), ("STK0", 400), ("STK1", 390), ("BIG", 5000), ("MID", 2000), ("ZERO", 100)],
    rmsk=ANNOT1 + _stack + [
        rmsk_row("chr1", 999900, 1000000, "+", "AluY", "SINE", "Alu", 1, 101, -199),          # touches chromosome end
        rmsk_row("chr3", 1000000, 3500000, "+", "BIG", "Satellite", "centr", 1, 2500001, 0),  # level >= 3 bin
        rmsk_row("chr3", 1048000, 1049200, "-", "MID", "LTR", "ERV1", -10, 1200, 1),          # crosses 1<<20
        rmsk_row("chr3", 1048570, 1048580, "+", "AluY", "OTHERCLASS", "OTHERFAM", 1, 11, 0),  # differs from first AluY row
        rmsk_row("chr3", 2000000, 2000000, "+", "ZERO", "Low", "Low", 1, 1, 0),               # zero length
        rmsk_row("chrUnknown", 10, 500, "+", "AluY", "SINE", "Alu", 1, 300, 0),              # chromosome absent: dropped
        rmsk_row("chrDup", 100000, 100300, "+", "AluY", "SINE", "Alu", 1, 300, 0),           # uses the LAST chrDup size
        rmsk_row("chr1", 30000, 30300, "+", "NoSize", "SINE", "Alu", 0x10, 0o377, 0)],         # strtol base 0
    refs=[("chr1", 1000000), ("chr2", 2), ("chr3", 5000000), ("chrNope", 1000), ("chrDup", 300000)],
    reads=[se("e1", 0, 999950, 40), se("e2", 16, 999970, 40), se("e3", 0, 999999, 40, cigar="1M"),
           se("n1", 0, 10, 40, tid=3), se("n2", 16, 20, 40, tid=3), se("s2", 0, 0, 40, tid=1, cigar="1M"),
           se("k1", 0, 20100, 40), se("k2", 16, 20180, 3), se("k3", 0, 20010, 40, cigar="20M100N16M"),
           se("b1", 0, 1048400, 40, tid=2), se("b2", 16, 1048560, 12, tid=2), se("b3", 0, 2000000 - 10, 40, tid=2),
           se("b4", 0, 1048565, 40, tid=2, cigar="30=6X"), se("d1", 0, 1050, 37), se("d1b", 0, 1050, 37),
           se("d2", 0, 1050, 3), se("d3", 0, 1050, 3), se("u1", 0, 100100, 40, tid=4), se("m1", 16, 20, 40, cigar="36M"),
           se("h1", 0, 30010, 40), se("z1", 0, 5, 40, cigar="*")],
    variants={"default": ("stat", ["-w"]), "E0": ("stat", ["-w", "-E", "0"]), "R": ("stat", ["-w", "-R"]),
              "E1000": ("stat", ["-w", "-E", "1000"]), "c09": ("stat", ["-w", "-c", "0.9"]),
              "filter_r": ("filter", ["-r"]), "filter_R": ("filter", ["-R", "-r"])})

# -C: reference names without the chr prefix, MT, GL contigs.
KATS["kat7_addchr"] = dict(
    chrom=CHR1M + [("chrM", 16571)], rep=REP1,
    rmsk=ANNOT1 + [rmsk_row("chrM", 100, 400, "+", "AluY", "SINE", "Alu", 1, 300, 0)],
    refs=[("1", 1000000), ("MT", 16571), ("GL000207.1", 4262), ("chr1", 1000000), ("mt", 16571)],
    reads=[se("a1", 0, 1050, 37, tid=0), se("a2", 0, 150, 37, tid=1), se("a3", 0, 10, 37, tid=2), se("a4", 16, 1200, 37, tid=3),
           se("a5", 0, 120, 2, tid=4)],
    variants={"C": ("stat", ["-w", "-C"]), "noC": ("stat", ["-w"]), "filter_C": ("filter", ["-C", "-r"])})

# aux scan quirks (bam_aux.c:28-34 with bam.h:754-760): 'f' values are skipped with size 0, so tags
# behind a float are mis-parsed; B arrays; NM as C / s / absent.
KATS["kat8_aux"] = dict(
    chrom=CHR1M, rep=REP1, rmsk=ANNOT1, refs=CHR1M,
    reads=[se("x1", 0, 1050, 0, aux=[("NM", "C", 1), ("XA", "Z", "chr1,+5101,36M,1;")]),
           se("x2", 0, 1050, 0, aux=[("NM", "s", 1), ("XA", "Z", "chr1,-5101,36M,1;chr1,+1101,36M,0;")]),
           se("x3", 0, 1050, 0, aux=[("XA", "Z", "chr1,+5101,36M,0;")]),
           se("x4", 0, 1050, 0, aux=[("XA", "Z", "chr1,+5101,36M,1;")]),
           se("x5", 0, 1050, 0, aux=[("AS", "f", 1.5), ("NM", "i", 1), ("XA", "Z", "chr1,+5101,36M,1;")]),
           se("x6", 0, 1050, 0, aux=[("ZB", "B", b"S" + (3).to_bytes(4, "little") + bytes(6)), ("NM", "i", 2), ("XA", "Z", "chr1,+5101,36M,2;chrQ,+5,36M,0;")]),
           se("x7", 0, 1050, 0, aux=[("NM", "i", 0), ("XA", "Z", "chr1,+1201,36M,0;")]),
           se("x8", 0, 5910, 0, aux=[("NM", "i", 0), ("XA", "Z", "chr1,+1201,36M,0;")]),
           se("x9", 0, 1050, 0, aux=[("NM", "i", 3), ("XA", "Z", "chr1,+5990,36M,3;")]),
           se("x10", 16, 1250, 0, aux=[("X0", "i", 2), ("X1", "A", "c"), ("MD", "Z", "36"), ("NM", "c", -1), ("XA", "Z", "chr1,+5101,36M,0;")])],
    variants={"default": ("stat", ["-w"]), "x": ("stat", ["-w", "-x"]), "B": ("stat", ["-w", "-B"])})
