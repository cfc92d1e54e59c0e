"""The score parser k_bedgraph runs on the device (itx_strtod_fast, host build) against the C library's strtod as Python
exposes it: wherever it claims an exact answer the bits must be the same; everything else must be handed to the host."""
import random
import struct

import emu_lib


def bits(x):
    return struct.pack("<d", x)


_LC = None


def libc(text):
    # strtod semantics: longest valid prefix
    global _LC
    if _LC is None:
        import ctypes, ctypes.util
        _LC = ctypes.CDLL(ctypes.util.find_library("c"))
        _LC.strtod.restype = ctypes.c_double
        _LC.strtod.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    return _LC.strtod(text.encode(), None)


def test_exact_where_claimed_and_honest_elsewhere():
    rnd = random.Random(3)
    cases = ["0", "0.0", "-0", "1", "12.34", "49.99", "0.01", ".5", "5.", "+3.25", "-7.125", "1e3", "1E-3", "2.5e+2", "1e22", "1e23", "1e-22",
             "1e-23", "9007199254740991", "9007199254740992", "9007199254740993", "123456789012345678", "0.1234567890123456789", "1.0000000000000000000001",
             "12abc", "1e", "1e+", "3.5.7", "inf", "-inf", "nan", "infinity", "0x1p3", "0x10", "", "abc", "  7.5", "\t2.25", "00012.5000", "1e400", "1e-400",
             "123456789012345678901234567890", "0.000000000000000000000000000001", "4.9e-324", "1.7976931348623157e308"]
    for _ in range(20000):
        d = rnd.choice([1, 2, 3, 6, 10, 15, 17, 19, 22])
        s = "".join(rnd.choice("0123456789") for _ in range(d))
        k = rnd.randrange(0, len(s) + 1)
        t = (rnd.choice(["", "-", "+"]) + s[:k] + rnd.choice(["", "."]) + s[k:] + rnd.choice(["", "", "", "e%d" % rnd.randrange(-30, 30)]))
        cases.append(t)
    n_exact = 0
    for t in cases:
        v, exact = emu_lib.strtod_fast(t)
        if exact:
            n_exact += 1
            assert bits(v) == bits(libc(t)) or (v == 0.0 and libc(t) == 0.0), t
    assert n_exact > len(cases) // 2
    for t in ("inf", "nan", "0x1p3", "9007199254740993", "1e23", "0.1234567890123456789", "abc", ""):
        assert not emu_lib.strtod_fast(t)[1], t
    for t in ("12.34", "49.99", "1e22", "9007199254740991", "  7.5", "12abc"):
        assert emu_lib.strtod_fast(t)[1], t
