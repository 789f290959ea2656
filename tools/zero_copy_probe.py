#!/usr/bin/env python3
"""Which operands pay when the host-pointer calls use pinned buffers in place:
times encrypt and decrypt of one 65536 batch (2048-bit key) for every value of
the IPCLB200_ZERO_COPY mask (1 = plaintexts in, 2 = ciphertexts out, 4 =
ciphertexts in of decrypt).  One JSON line per mask."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)["2048"].items()}
    p, q = sorted((k["p"], k["q"]))
    capi.init(0)
    os.environ["IPCLB200_COMB_SYNC"] = "1"
    pk = capi.PubKey(to_limbs(p * q, 64), to_limbs(k["hs"], 128), 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    rng = np.random.default_rng(5)
    pt = capi.pinned_empty((count, 64))
    r = capi.pinned_empty((count, 32))
    ct = capi.pinned_empty((count, 128))
    dt = capi.pinned_empty((count, 64))
    pt[:] = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    r[:] = random_limbs(rng, count, 32)
    for _ in range(3):
        pk.encrypt(pt, r, out=ct)
        sk.decrypt(ct, out=dt)
    for mask in (0, 1, 2, 4, 3, 7):
        os.environ["IPCLB200_ZERO_COPY"] = str(mask)
        enc, dec = [], []
        for _ in range(6):
            t0 = time.perf_counter()
            pk.encrypt(pt, r, out=ct)
            t1 = time.perf_counter()
            sk.decrypt(ct, out=dt)
            t2 = time.perf_counter()
            enc.append((t1 - t0) * 1e3)
            dec.append((t2 - t1) * 1e3)
        assert np.array_equal(dt, pt)
        print(json.dumps({"count": count, "zero_copy_mask": mask,
                          "encrypt_ms": round(float(np.median(enc)), 3),
                          "decrypt_ms": round(float(np.median(dec)), 3)}), flush=True)


if __name__ == "__main__":
    main()
