"""Tiny pass through every kernel, meant to run under compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Results are checked against Python integers."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,  # noqa: E402
                                          random_limbs, to_limbs)

capi.init(0)
rng = np.random.default_rng(11)
with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
    K = {b: {k: int(v, 16) for k, v in d.items()} for b, d in json.load(f).items()}

for L in (16, 32, 48, 64, 96, 128):
    count = 9
    mod = random_limbs(rng, 1, L)
    mod[0, 0] |= 1
    base, exp = random_limbs(rng, count, L), random_limbs(rng, count, 2)
    got = batch_from_limbs(capi.modexp(base, exp, mod, capi.SHARED_MOD))
    m = batch_from_limbs(mod)[0]
    want = [pow(b, e, m) for b, e in zip(batch_from_limbs(base), batch_from_limbs(exp))]
    assert got == want, L
    a, b = random_limbs(rng, count, L), random_limbs(rng, count, L)
    got = batch_from_limbs(capi.modmul(a, b, mod[0]))
    assert got == [x * y % m for x, y in zip(batch_from_limbs(a), batch_from_limbs(b))]

for bits in ("1024", "2048"):
    k = K[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    pk0 = capi.PubKey(to_limbs(n, NL))
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    for count, no_comb in ((5, "1"), (70, "0")):   # windowed and comb obfuscator
        pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
        r = random_limbs(rng, count, 2)      # short randoms keep it quick
        os.environ["IPCLB200_NO_COMB"] = no_comb
        ct = pk.encrypt(pt, r)
        del os.environ["IPCLB200_NO_COMB"]
        want = [(n * m + 1) * pow(k["hs"], e, n * n) % (n * n)
                for m, e in zip(batch_from_limbs(pt), batch_from_limbs(r))]
        assert batch_from_limbs(ct) == want
    pt = random_limbs(rng, 6, NL, top_mask=0x3FFFFFFF)
    ct = pk0.encrypt(pt, random_limbs(rng, 6, 1))
    assert np.array_equal(sk.decrypt(ct), pt)
    assert np.array_equal(sk.decrypt(ct, use_crt=False), pt)
    os.environ["IPCLB200_DECRYPT"] = "tile"
    assert np.array_equal(sk.decrypt(ct), pt)
    del os.environ["IPCLB200_DECRYPT"]
    assert np.array_equal(sk.decrypt(pk0.encrypt(pt, None, make_secure=False)), pt)

# ---- round-2 kernels: every lane layout of the two-digit CRT decrypt (incl. the
# thread-per-task kernel with its TMA fetch into the compact staging area), the
# constant schedule, the two-digit ct*pt ladder, device-drawn randoms
k = K["2048"]
p, q = sorted((k["p"], k["q"]))
n = p * q
pk = capi.PubKey(to_limbs(n, 64), to_limbs(k["hs"], 128), 1024)
sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
pt = random_limbs(rng, 37, 64, top_mask=0x3FFFFFFF)
ct = pk.encrypt(pt, random_limbs(rng, 37, 2))
for layout in ("-2", "-1", "0", "1", "2"):
    os.environ["IPCLB200_HENSEL_SPREAD"] = layout
    os.environ["IPCLB200_HENSEL_W64"] = "1"
    assert np.array_equal(sk.decrypt(ct), pt), layout
sk.set_schedule(True)
os.environ["IPCLB200_HENSEL_SPREAD"] = "-2"
assert np.array_equal(sk.decrypt(ct[:5]), pt[:5])
del os.environ["IPCLB200_HENSEL_SPREAD"], os.environ["IPCLB200_HENSEL_W64"]
nsq = to_limbs(n * n, 128)
e = random_limbs(rng, 7, 2)
got = batch_from_limbs(capi.modexp(ct[:7], e, nsq, capi.SHARED_MOD))
assert got == [pow(c, x, n * n) for c, x in zip(batch_from_limbs(ct[:7]), batch_from_limbs(e))]
key = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32)
nonce = rng.integers(0, 2**32, 3, dtype=np.uint64).astype(np.uint32)
c2 = pk.encrypt_drbg(pt[:9], key, nonce)
assert np.array_equal(sk.decrypt(c2), pt[:9])
b = capi.Batch(33, 18)
b.random(561, key, nonce)
assert b.download().shape == (33, 18)
capi.shutdown()
print("sanitize_smoke ok")
