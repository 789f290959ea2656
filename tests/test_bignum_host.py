"""Host-only checks of the from-scratch ::BigNumber (pailliercryptolib_b200/
ipcl/src/bignum.cpp) against Python integers, through tests/cpp/bn_shim.cpp.
Semantics checked are the reference's (ipcl/bignum.cpp): non-negative %,
"0x" lower-case hex without leading zeros, one word for zero, big-endian
toBin/fromBin."""
import ctypes
import random

import pytest


@pytest.fixture(scope="module")
def bn():
    from pailliercryptolib_b200 import build
    build.build_cpp_tests()
    lib = ctypes.CDLL(build.BN_SHIM)
    buf = ctypes.create_string_buffer(1 << 16)

    def call(op, a=0, b=0, c=0, raw=False):
        def s(x):
            if isinstance(x, str):
                return x.encode()
            return (("-" if x < 0 else "") + hex(abs(x))).encode()
        rc = lib.bn_op(op.encode(), s(a), s(b), s(c), buf, len(buf))
        out = buf.value.decode()
        if rc != 0:
            raise RuntimeError(out)
        if raw:
            return out
        neg = out.startswith("-")
        body = out[3:] if neg else out[2:]
        v = int(body, 16) if body else 0
        return -v if neg else v
    return call


def rand_int(rnd, bits, signed=True):
    v = rnd.getrandbits(rnd.choice([1, 31, 32, 33, 64, 100, bits]))
    return -v if signed and rnd.random() < 0.3 else v


def test_arithmetic_random(bn):
    rnd = random.Random(1)
    for _ in range(400):
        a, b = rand_int(rnd, 2048), rand_int(rnd, 1024)
        assert bn("add", a, b) == a + b
        assert bn("sub", a, b) == a - b
        assert bn("mul", a, b) == a * b
        assert bn("cmp", a, b, raw=True) == str((a > b) - (a < b))
        if b != 0:
            q = abs(a) // abs(b)
            q = -q if (a < 0) != (b < 0) else q
            assert bn("div", a, b) == q               # truncated toward zero
        if b > 0:
            assert bn("mod", a, b) == a % b           # non-negative residue
        w = rnd.getrandbits(32)
        assert bn("addw", a, str(w)) == a + w
        assert bn("mulw", a, str(w)) == a * w


def test_knuth_division_corner_cases(bn):
    B = 1 << 32
    cases = [
        ((B ** 4) - 1, (B ** 2) - 1), (B ** 5, B ** 2 + 1), ((B ** 3) * 0x7fffffff, (B ** 2) * 0x80000000 + 1),
        (0x8000000000000000_0000000000000000, 0x8000000000000001),
        ((1 << 4096) - 1, (1 << 2048) - 1), ((1 << 4096), (1 << 2047) + 1),
        (12345, 1 << 200), (0, 7), ((B - 1) * B ** 3 + (B - 1), (B - 1) * B + (B - 1)),
    ]
    rnd = random.Random(5)
    for _ in range(300):
        nb = rnd.choice([2, 3, 4, 8, 64])
        d = rnd.getrandbits(32 * nb) | (1 << (32 * nb - rnd.choice([1, 2, 17, 32])))
        a = d * rnd.getrandbits(32 * rnd.choice([1, 2, nb])) + rnd.randrange(d)
        cases.append((a, d))
    for a, d in cases:
        assert bn("div", a, d) == a // d
        assert bn("mod", a, d) == a % d


def test_modular_helpers(bn):
    rnd = random.Random(2)
    import math
    for _ in range(100):
        m = rnd.getrandbits(512) | (1 << 511) | 1
        a, b = rnd.getrandbits(600), rnd.getrandbits(500)
        assert bn("modmul", a, b, m) == a * b % m
        assert bn("modadd", a, b, m) == (a + b) % m
        assert bn("modsub", a, b, m) == (a - b) % m
        assert bn("invadd", a, m) == (-a) % m
        assert bn("gcd", a, b) == math.gcd(a, b)
        if math.gcd(a, m) == 1:
            assert bn("invmul", a, m) == pow(a, -1, m)
    with pytest.raises(RuntimeError):
        bn("invmul", 6, 9)
    with pytest.raises(RuntimeError):
        bn("div", 5, 0)


def test_string_formats(bn):
    assert bn("dec", "12345678901234567890123", raw=True) == hex(12345678901234567890123)
    assert bn("dec", "-255", raw=True) == "-0xff"
    assert bn("dec", "0XABCDEF", raw=True) == "0xabcdef"
    assert bn("dec", "0x000012", raw=True) == "0x12"
    # zero prints as "0x" and is one word / one bit wide (ippsRef_BN convention)
    assert bn("dec", "0", raw=True) == "0x"
    assert bn("vec", 0, raw=True) == "0"
    assert bn("vec", (5 << 32) | 7, raw=True) == "7,5"
    assert bn("bits", 0, "0", raw=True) == "1,0,0,1,0,0"
    assert bn("bits", 0b101000, "3", raw=True) == "6,5,3,1,0,1"
    assert bn("bits", (1 << 64) | 1, "64", raw=True) == "65,64,0,3,1,1"


def test_bin_roundtrip_big_endian(bn):
    v = 0x0102030405060708090a0b0c
    out = bn("bin", v, "16", raw=True)
    octets, back = out.split(",")
    assert octets == "000000000102030405060708090a0b0c"
    assert int(back, 16) == v
