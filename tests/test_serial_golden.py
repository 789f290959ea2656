"""The committed golden archives of ipcl::serializer are exactly what the
generator derives from cereal's PortableBinary rules (guards against drift of
tests/golden/serial_golden.json and tests/cpp/serial_golden.hpp); the C++ side
loads and re-saves them byte for byte in tests/cpp (SerialTest.GoldenArchives)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_reproduces_committed_goldens(tmp_path):
    gold = os.path.join(ROOT, "tests", "golden", "serial_golden.json")
    hdr = os.path.join(ROOT, "tests", "cpp", "serial_golden.hpp")
    before = open(gold).read(), open(hdr).read()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden",
                                                     "make_serial_golden.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert (open(gold).read(), open(hdr).read()) == before


def test_golden_bignum_layout_by_hand():
    """the smallest archive spelled out: endian flag, class version, u64 word
    count, words LSW first, sign"""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "serial_golden.json")))
    b = bytes.fromhex(g["archives"]["bignum_pos"])
    assert b[0] == 1 and b[1:5] == b"\0\0\0\0"
    assert int.from_bytes(b[5:13], "little") == 4
    words = [int.from_bytes(b[13 + 4 * i:17 + 4 * i], "little") for i in range(4)]
    assert sum(w << (32 * i) for i, w in enumerate(words)) == 0x1234567890ABCDEF0011223344556677
    assert int.from_bytes(b[29:33], "little") == 1
    n = bytes.fromhex(g["archives"]["bignum_neg"])
    assert int.from_bytes(n[-4:], "little") == 0
