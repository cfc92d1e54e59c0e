"""The device DEFLATE decoder (iteres_b200/csrc/itx_inflate.cuh, host build) against zlib: stored, fixed and
dynamic blocks, every compression level, several kinds of content incl. the BAM streams the generator makes,
and damaged input (must report an error, never run away)."""
import random
import struct
import zlib

import numpy as np
import pytest

import emu_lib
import synth


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(data) + co.flush()


def samples():
    rnd = random.Random(4)
    out = {
        "empty": b"",
        "one": b"A",
        "zeros": bytes(65280),
        "run": b"ab" * 30000,
        "random": bytes(rnd.getrandbits(8) for _ in range(40000)),
        "text": (b"the quick brown fox jumps over the lazy dog, " * 1400)[:65000],
        "lowent": bytes(rnd.choice(b"ACGT") for _ in range(65280)),
        "ramp": bytes(range(256)) * 250,
        "short_periods": b"".join(bytes([65 + i % 26]) * (3 + i % 7) + b"xyz" * (2 + i % 5) + b"abcdefg" * 3 for i in range(2500))[:65000],
        "triples": bytes(rnd.choice(b"ab") for _ in range(65000)),       # dense 3..6-byte matches: long match lists
        # the largest block BGZF allows (ISIZE 65536: positions fill 16 bits, the last window of the match copies is cut short)
        "max_block": (b"".join(bytes([65 + i % 26]) * (3 + i % 7) + b"xyz" * (2 + i % 5) + bytes(rnd.randrange(256) for _ in range(i % 11)) for i in range(4000)))[:65536],
    }
    s = synth.Synth(0, 2000, seed=5)
    for mode in (0, 1, 2):
        buf, n, _ = s.stream(mode, 3000)
        out["bam%d" % mode] = buf[:n].tobytes()[: 0xff00]
        out["bam%d_tail" % mode] = buf[:n].tobytes()[0xff00: 2 * 0xff00]
    s.close()
    return out


SAMPLES = samples()


@pytest.mark.parametrize("defer", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("name", sorted(SAMPLES))
def test_matches_zlib(name, level, defer):
    data = SAMPLES[name]
    rc, got = emu_lib.inflate(deflate(data, level), len(data), defer=defer)
    assert rc == 0 and got == data


@pytest.mark.parametrize("name", ["text", "bam0", "lowent", "one"])
def test_fixed_huffman_and_other_strategies(name):
    data = SAMPLES[name]
    for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        rc, got = emu_lib.inflate(deflate(data, 6, strat), len(data))
        assert rc == 0 and got == data, strat


def test_multi_block_stream_with_sync_flushes():
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = SAMPLES["text"]
    comp = b""
    for i in range(0, len(data), 7000):
        comp += co.compress(data[i:i + 7000]) + co.flush(zlib.Z_SYNC_FLUSH)      # stored empty blocks in between
    comp += co.flush()
    rc, got = emu_lib.inflate(comp, len(data))
    assert rc == 0 and got == data


def test_wrong_isize_and_damaged_streams_are_errors():
    data = SAMPLES["bam0"]
    comp = deflate(data, 6)
    assert emu_lib.inflate(comp, len(data) + 1, cap=len(data) + 1)[0] == 2           # ISIZE mismatch
    assert emu_lib.inflate(comp, len(data) - 5, cap=len(data) - 5)[0] != 0           # would overflow the block: refused
    assert emu_lib.inflate(comp[: len(comp) // 2], len(data))[0] != 0                # truncated
    rnd = random.Random(9)
    bad = 0
    for _ in range(200):
        b = bytearray(comp)
        for _ in range(3):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        rc, got = emu_lib.inflate(bytes(b), len(data))
        if rc != 0 or got != data:
            bad += 1
    assert bad > 150                                                                   # and none of them hung or crashed
    for _ in range(100):                                                               # pure noise
        junk = bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(1, 300)))
        emu_lib.inflate(junk, 1000)


def test_bgzf_blocks_of_a_generated_bam(tmp_path):
    s = synth.Synth(1, 5000, seed=8)
    bam = str(tmp_path / "x.bam")
    n, nrec = s.write_bam(bam, 2, 4000, level=1, threads=2)
    buf, n2, _ = s.stream(2, 4000)
    raw = open(bam, "rb").read()
    off, out = 0, b""
    while off + 18 <= len(raw):
        bsize = struct.unpack_from("<H", raw, off + 16)[0] + 1
        isize = struct.unpack_from("<I", raw, off + bsize - 4)[0]
        rc, got = emu_lib.inflate(raw[off + 18: off + bsize - 8], isize)
        assert rc == 0
        out += got
        off += bsize
    assert out == buf[:n2].tobytes()
    s.close()
