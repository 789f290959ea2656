"""The bit-level Python models of the device arithmetic (tools/model_*.py) are
the CPU-checkable specification of the kernels: same limb layouts, same carry
handling, word-size bounds asserted.  These tests run them against Python
integers, so an index or bound mistake in a redesign shows up without a GPU."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_cios_lane_model():
    import model_montmul as m
    assert hasattr(m, "row")
    # the module's own self-test entry point, if it has one
    for name in ("main", "self_test"):
        fn = getattr(m, name, None)
        if fn:
            fn()
            break


def test_fp64_montgomery_model():
    import model_fp64_mont as m
    rnd = random.Random(5)
    K, T = 24, 4
    L = K * T
    for trial in range(4):
        n = rnd.getrandbits(2048) | (1 << 2047) | 1
        R = 1 << (m.W * L)
        n0inv = (-pow(n, -1, 1 << m.W)) % (1 << m.W)
        a, b = (2 * n - 1, 2 * n - 1) if trial == 0 else (rnd.randrange(2 * n), rnd.randrange(2 * n))
        r = m.mont_mul(m.split(m.to_limbs(a, L), K, T), m.to_limbs(b, L),
                       m.split(m.to_limbs(n, L), K, T), n0inv, K, T)
        got = m.from_limbs(m.join(r))
        assert got % n == a * b * pow(R, -1, n) % n and got <= n + 1
    p = rnd.getrandbits(1024) | (1 << 1023) | 1
    ct = rnd.getrandbits(4090)
    assert m.decrypt_side(ct, p) == pow(ct, p - 1, p * p)


def test_symmetric_squaring_models():
    import model_sqr as m
    rnd = random.Random(9)
    for a in [(1 << 2048) - 1, 0, 1, int("ffffffff00000000" * 32, 16)] + \
            [rnd.getrandbits(2048) for _ in range(6)]:
        W, H, _ = m.sqr_product(a)          # 16 x 4 layout
        assert m.val(W) + (m.val(H) << 2048) == a * a
        assert m.val(m.sqr2_product(a)) == a * a   # 32 x 2 layout
        n = rnd.getrandbits(2048) | (1 << 2047) | 1
        r = m.mont_sqr(a, n)
        assert r < (1 << 2048) and (r * (1 << 2048) - a * a) % n == 0
