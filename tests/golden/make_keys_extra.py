"""Extra key fixtures for the two-digit (Hensel) decrypt kernel
(tests/golden/keys_extra.json): a 4096-bit key (2048-bit primes: the 16 x 4
layout) and two 2048-bit-class keys whose primes sit at the two ends of the
range the kernel accepts -- just above R/2 (the most conditional subtractions
per product) and just below R = 2^1024.  Seeded, Python only; the expected
plaintexts in the tests come from Python pow() as in make_golden.py.

    python tests/golden/make_keys_extra.py
"""
import json
import math
import os
import random

from make_golden import djn_hs, djn_keypair, is_prime

HERE = os.path.dirname(os.path.abspath(__file__))


def prime_near(start, step, rnd):
    c = start | 3
    while not (is_prime(c, rnd) and c % 4 == 3):
        c += step
    return c


def main():
    rnd = random.Random(0xB200E)
    out = {}
    p, q = djn_keypair(4096, rnd)
    out["4096"] = dict(p=p, q=q)
    lo = 1 << 1023
    while True:
        p = prime_near(lo + rnd.getrandbits(64) * 4, 4, rnd)
        q = prime_near(lo + rnd.getrandbits(64) * 4, 4, rnd)
        if p != q and math.gcd(p - 1, q - 1) == 2:
            break
    out["2048_low"] = dict(p=p, q=q)
    hi = (1 << 1024) - 1
    while True:
        p = prime_near(hi - rnd.getrandbits(64) * 4 - 4, -4, rnd)
        q = prime_near(hi - rnd.getrandbits(64) * 4 - 4, -4, rnd)
        if p != q and math.gcd(p - 1, q - 1) == 2:
            break
    out["2048_high"] = dict(p=p, q=q)
    for k in out.values():
        k["hs"] = djn_hs(k["p"] * k["q"], rnd)
    with open(os.path.join(HERE, "keys_extra.json"), "w") as f:
        json.dump({name: {a: format(b, "x") for a, b in k.items()}
                   for name, k in out.items()}, f, indent=1)
    print({name: (k["p"].bit_length(), k["q"].bit_length()) for name, k in out.items()})


if __name__ == "__main__":
    main()
