"""Minimal BAM/BGZF writer for tests (pure Python, zlib).  Records are given as dicts; the encoder
follows the SAM/BAM specification (little-endian core of 8 x u32, CIGAR op codes MIDNSHP=X)."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def parse_cigar(s):
    out, num = [], ""
    if s in ("*", ""):
        return out
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | CIGAR_OPS.index(ch))
            num = ""
    return out


def encode_aux(aux):
    """aux: list of (tag, type, value); type in A c C s S i I f Z B(raw bytes incl. subtype+count)."""
    b = b""
    for tag, ty, val in aux:
        b += tag.encode() + ty.encode()
        if ty == "A":
            b += val.encode()[:1]
        elif ty in "cCsSiIf":
            b += struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[ty], val)
        elif ty in "ZH":
            b += val.encode() + b"\0"
        elif ty == "B":
            b += val
    return b


def encode_record(r):
    qname = r["qname"].encode() + b"\0"
    cig = parse_cigar(r.get("cigar", "*"))
    seq = r.get("seq", "*")
    l_seq = 0 if seq == "*" else len(seq)
    pos = r.get("pos", -1)
    end = pos
    for c in cig:
        if (c & 0xF) in (0, 2, 3):
            end += c >> 4
    if end == pos:
        end = pos + 1
    bin_ = reg2bin(pos, end) if pos >= 0 else 4680
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    sb = bytearray((l_seq + 1) // 2)
    for i in range(l_seq):
        v = code.get(seq[i].upper(), 15)
        sb[i >> 1] |= v << 4 if i % 2 == 0 else v
    qual = r.get("qual", "*")
    qb = bytes([0xFF] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    aux = encode_aux(r.get("aux", []))
    core = struct.pack("<iiIIiiii", r.get("tid", -1), pos,
                       (bin_ << 16) | (r.get("mapq", 0) << 8) | len(qname),
                       (r.get("flag", 0) << 16) | len(cig), l_seq,
                       r.get("mtid", -1), r.get("mpos", -1), r.get("isize", 0))
    data = core + qname + b"".join(struct.pack("<I", c) for c in cig) + bytes(sb) + qb + aux
    return struct.pack("<i", len(data)) + data


def encode_header(refs, text=None):
    """refs: list of (name, length)."""
    if text is None:
        text = "@HD\tVN:1.0\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in refs)
    t = text.encode()
    b = b"BAM\1" + struct.pack("<i", len(t)) + t + struct.pack("<i", len(refs))
    for n, l in refs:
        nb = n.encode() + b"\0"
        b += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    return b


def bgzf_block(data, level=6):
    if level == 0:
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
    else:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = 18 + len(comp) + 8 - 1
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


EOF_BLOCK = bgzf_block(b"")


def write_bam(path, refs, records, block=0xFF00, level=6, eof=True, text=None):
    """records: list of dicts (encoded here) or already-encoded bytes."""
    stream = encode_header(refs, text) + b"".join(r if isinstance(r, (bytes, bytearray)) else encode_record(r)
                                                    for r in records)
    with open(path, "wb") as f:
        for i in range(0, len(stream), block):
            f.write(bgzf_block(stream[i:i + block], level))
        if eof:
            f.write(EOF_BLOCK)
    return stream


def sam_line(r, refs):
    tid, mtid = r.get("tid", -1), r.get("mtid", -1)
    rn = "*" if tid < 0 else refs[tid][0]
    mr = "*" if mtid < 0 else ("=" if mtid == tid else refs[mtid][0])
    f = [r["qname"], str(r.get("flag", 0)), rn, str(r.get("pos", -1) + 1), str(r.get("mapq", 0)), r.get("cigar", "*"),
         mr, str(r.get("mpos", -1) + 1), str(r.get("isize", 0)), r.get("seq", "*"), r.get("qual", "*")]
    for tag, ty, val in r.get("aux", []):
        f.append("%s:%s:%s" % (tag, "i" if ty in "cCsSiI" else ty, val))
    return "\t".join(f)
