"""Randomised differential run on one GPU: batches of random size (1 ... 70000,
biased towards the boundaries of the layout cost model) at the golden keys go
through encrypt -> decrypt via the host-pointer C ABI with the library's own
layout choice; every plaintext must come back, the ciphertexts of a random slice
are compared with the oracle (AVX512-IFMA mb8 restatement at 2048 bits, scalar
otherwise), ct+ct and ct*pt of a slice with Python integers.  One JSON line per
batch.  `python tools/fuzz_roundtrip.py [seconds] [seed]`; run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pailliercryptolib_b200 import build, capi  # noqa: E402
from pailliercryptolib_b200.limbs import (batch_from_limbs, random_limbs,  # noqa: E402
                                          to_limbs)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    build.build_oracle()
    import oracle as orc
    keys = {}
    for name in ("keys.json", "keys_extra.json"):
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            keys.update({b: {k: int(v, 16) for k, v in d.items()} for b, d in json.load(f).items()})
    capi.init(0)
    rng = np.random.default_rng(seed)
    edges = [1, 2, 15, 16, 17, 31, 33, 511, 1024, 2047, 2049, 4096, 7105, 8192, 14207, 14209,
             14336, 20480, 28415, 28417, 32768, 45055, 45057, 56833, 65536, 70000]
    objs = {}
    for bits in ("1024", "2048", "3072", "4096", "2048_low", "2048_high"):
        k = keys[bits]
        p, q = sorted((k["p"], k["q"]))
        NL = int(bits.split("_")[0]) // 32
        objs[bits] = (p, q, NL, capi.PubKey(to_limbs(p * q, NL), to_limbs(k["hs"], 2 * NL),
                                            NL * 16),
                      capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2)), k["hs"],
                      capi.PubKey(to_limbs(p * q, NL)))
    t_end = time.time() + budget
    n_batches = 0
    while time.time() < t_end:
        bits = str(rng.choice(["1024", "2048", "2048", "2048", "3072", "4096", "2048_low",
                               "2048_high"]))
        p, q, NL, pk, sk, hs, pk_std = objs[bits]
        n = p * q
        if rng.random() < 0.6:
            count = int(rng.choice(edges)) + int(rng.integers(-1, 2))
        else:
            count = int(rng.integers(1, 70001))
        count = max(1, count)
        if bits in ("3072", "4096"):
            count = min(count, 20000 if bits == "3072" else 6000)
        pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
        r = random_limbs(rng, count, NL // 2)
        ct = pk.encrypt(pt, r)
        got = sk.decrypt(ct)
        ok = bool(np.array_equal(got, pt))
        lo = int(rng.integers(0, count))
        hi = min(count, lo + 64)
        nl, hsl = to_limbs(n, NL), to_limbs(hs, 2 * NL)
        if bits.startswith("2048") and orc.have_ifma():
            want = orc.encrypt_mb8(nl, hsl, pt[lo:hi], r[lo:hi])
        else:
            want = orc.encrypt(nl, hsl, pt[lo:hi], r[lo:hi])
        ok_ct = bool(np.array_equal(ct[lo:hi], want))
        # homomorphic ops on the slice
        nsq = n * n
        a, b = batch_from_limbs(ct[lo:hi]), batch_from_limbs(ct[lo:hi][::-1])
        s = capi.modmul(ct[lo:hi], np.ascontiguousarray(ct[lo:hi][::-1]), to_limbs(nsq, 2 * NL))
        ok_add = batch_from_limbs(s) == [x * y % nsq for x, y in zip(a, b)]
        e = random_limbs(rng, hi - lo, 1)
        m = capi.modexp(ct[lo:hi], e, to_limbs(nsq, 2 * NL), capi.SHARED_MOD)
        ok_mul = batch_from_limbs(m) == [pow(x, int(y), nsq) for x, y in
                                         zip(a, batch_from_limbs(e))]
        # now and then the non-DJN obfuscator r^n and the RAW decrypt (both take the
        # two-digit ladder from 2048 elements on)
        ok_std = True
        if rng.random() < 0.15 and NL <= 64:
            c2 = min(count, 3000)
            rs = random_limbs(rng, c2, NL, top_mask=0x3FFFFFFF)
            rs[:, 0] |= 1
            cs = pk_std.encrypt(pt[:c2], rs)
            ok_std = bool(np.array_equal(sk.decrypt(cs, use_crt=False), pt[:c2]))
            ok_std = ok_std and bool(np.array_equal(
                cs[:8], orc.encrypt(nl, None, pt[:8], rs[:8])))
        layout = capi.decrypt_layout(count, NL // 2, 148)
        print(json.dumps({"bits": bits, "count": count, "layout": layout, "roundtrip": ok,
                          "ct_vs_oracle": ok_ct, "add": ok_add, "mul": ok_mul, "non_djn_raw": ok_std}),
              flush=True)
        n_batches += 1
        if not (ok and ok_ct and ok_add and ok_mul and ok_std):
            print(json.dumps({"FAILED": True}))
            sys.exit(1)
    print(json.dumps({"batches": n_batches, "all_ok": True}))


if __name__ == "__main__":
    main()
