"""Integer-level model of the two-digit (Hensel) arithmetic mod p^2 used by the
CRT-decrypt kernel (pailliercryptolib_b200/csrc/mont_hensel.cuh).

A residue X mod p^2 is held as a pair (x0, w), 0 <= x0, w < R = 2^(32*Lh),
meaning  X~ = x0 - w*p (mod p^2)  with X~ = X*R mod p^2 (Montgomery form).  A
product needs only HALF-WIDTH Montgomery passes mod p, because the p^2 term of
(x0 - wx*p)(y0 - wy*p) vanishes:

  pass A:  x0*y0 + m*p = TA*R          (CIOS mod p, quotient m = sum q_i 2^32i kept)
           z0 = TA - ovA*p             (ovA = [TA >= R])
  pass B:  U = x0*wy + wx*y0 + m ;  U + m'*p = V*R ;  wz = V - ovA  (mod p, < R)

so a multiply costs 5 half-width limb products (x0*y0, m*p, x0*wy, wx*y0, m'*p)
and a squaring 4 (2*x0*w is one product with the doubled multiplier), against 8
for the full-width Montgomery product mod p^2.  L(x) = (x-1)/p falls out of
digit 1 of x = c^(p-1): no exact division (see finish()).

This file checks the algebra and the bounds the kernel relies on against
Python's pow(); tools/model_hensel_words.py is the word-level model of the
accumulator rows.  Reference semantics: ipcl/pri_key.cpp:114-157."""
import random


class Hensel:
    def __init__(self, p, Lh):
        self.p = p
        self.Lh = Lh
        self.R = 1 << (32 * Lh)
        assert p % 2 == 1 and self.R // 2 <= p < self.R, "p must fill its limbs"
        self.pinv = (-pow(p, -1, self.R)) % self.R
        self.max_rounds = 0

    # ---- half-width Montgomery pass (what one CIOS sweep computes) ----------
    def redc_q(self, u):
        """u + m*p = t*R exactly; returns (t, m)."""
        m = (u * self.pinv) % self.R
        t = (u + m * self.p)
        assert t % self.R == 0
        return t // self.R, m

    def reduce_below_R(self, v):
        """rounds of 'if v >= R: v -= p' (warp-uniform loop in the kernel)"""
        rounds = 0
        while v >= self.R:
            v -= self.p
            rounds += 1
        self.max_rounds = max(self.max_rounds, rounds)
        assert rounds <= 3
        return v

    # ---- products ------------------------------------------------------------
    def hmul(self, x, y):
        x0, wx = x
        y0, wy = y
        R, p = self.R, self.p
        tA, m = self.redc_q(x0 * y0)
        assert tA < R + p
        ovA = 1 if tA >= R else 0
        z0 = tA - ovA * p
        U = x0 * wy + wx * y0 + m
        V, _ = self.redc_q(U)
        assert V < 2 * R + p
        assert V - ovA >= 0
        wz = self.reduce_below_R(V - ovA)
        return (z0, wz)

    def hsqr(self, x):
        x0, w = x
        R, p = self.R, self.p
        tA, m = self.redc_q(x0 * x0)
        ovA = 1 if tA >= R else 0
        z0 = tA - ovA * p
        dw = (2 * w) % R          # doubled multiplier, low Lh limbs
        hb = (2 * w) // R         # the bit that fell off: + x0*R before the division
        V, _ = self.redc_q(x0 * dw + m)
        V += hb * x0              # half row at the end of pass B
        assert V < 2 * R + p + 1
        assert V - ovA >= 0
        wz = self.reduce_below_R(V - ovA)
        return (z0, wz)

    def hadd(self, a, b):
        """(a0 - wa p) + (b0 - wb p); only used by the prologue."""
        R, p = self.R, self.p
        s0 = a[0] + b[0]
        subs = 0
        while s0 >= R:
            s0 -= p
            subs += 1
        assert subs <= 2
        # digit 0 went down by p  <=>  w goes down by 1  ==  + (p-1)
        w = a[1] + b[1] + subs * (p - 1)
        while w >= R:
            w -= p
        return (s0, w)

    def value(self, x):
        """the residue mod p^2 a pair stands for (out of Montgomery form)"""
        p2 = self.p * self.p
        return ((x[0] - x[1] * self.p) * pow(self.R, -1, p2)) % p2

    def const(self, v):
        """Hensel-Montgomery pair of an integer v (host side, key setup)"""
        p2 = self.p * self.p
        t = (v * self.R) % p2
        x0 = t % self.p
        w = (-(t // self.p)) % self.p
        assert (x0 - w * self.p) % p2 == t
        return (x0, w)

    # ---- prologue: ciphertext (4*Lh limbs) -> Montgomery pair ----------------
    def enter(self, c):
        R = self.R
        acc = None
        for j in range(4):
            cj = (c >> (32 * self.Lh * j)) % R
            kj = self.const(pow(R, j + 1, self.p * self.p))   # hmul divides by R once
            t = self.hmul((cj, 0), kj)
            acc = t if acc is None else self.hadd(acc, t)
        return acc

    # ---- epilogue: L(x) * hp mod p from digit 1 ------------------------------
    def finish(self, x, hp):
        """x = pair of c^(p-1) mod p^2 (== 1 mod p).  Returns (mp, regular):
        x0 must be R - p, then  x~ = R - (w+1) p = R + l R p  =>  l = -(w+1)/R,
        mp = l*hp = MontMul_p(w + 1, -hp)."""
        R, p = self.R, self.p
        regular = x[0] == R - p
        w1 = x[1] + 1
        if w1 == R:
            w1 -= p
        nhp = (-hp) % p
        t, _ = self.redc_q(w1 * nhp)
        return t % p, regular


def sliding_schedule(e, w=5):
    ops = []
    i = e.bit_length() - 1
    first = None
    while i >= 0:
        if not (e >> i) & 1:
            ops.append(0)
            i -= 1
            continue
        l = max(i - w + 1, 0)
        while not (e >> l) & 1:
            l += 1
        v = (e >> l) & ((1 << (i - l + 1)) - 1)
        if first is None:
            first = (v - 1) // 2
        else:
            ops += [0] * (i - l + 1) + [(v - 1) // 2 + 1]
        i = l - 1
    return first, ops


def hensel_pow(H, x, e):
    first, ops = sliding_schedule(e)
    x2 = H.hsqr(x)
    tab = [x]
    for _ in range(15):
        tab.append(H.hmul(tab[-1], x2))
    acc = tab[first]
    nsq = nmul = 0
    for o in ops:
        if o == 0:
            acc = H.hsqr(acc)
            nsq += 1
        else:
            acc = H.hmul(acc, tab[o - 1])
            nmul += 1
    return acc, nsq, nmul


def gen_prime(bits, rnd):
    while True:
        c = rnd.getrandbits(bits) | (1 << (bits - 1)) | 1
        if all(pow(a, c - 1, c) == 1 for a in (2, 3, 5, 7, 11, 13)):
            return c


def selftest(seed=7):
    rnd = random.Random(seed)
    for Lh in (2, 4, 16, 32):
        bits = 32 * Lh
        p = gen_prime(bits, rnd)
        q = gen_prime(bits, rnd)
        n = p * q
        H = Hensel(p, Lh)
        p2 = p * p
        # products of random pairs (not necessarily reduced below p)
        for _ in range(300):
            x = (rnd.randrange(H.R), rnd.randrange(H.R))
            y = (rnd.randrange(H.R), rnd.randrange(H.R))
            assert H.value(H.hmul(x, y)) == (H.value(x) * H.value(y)) % p2
            assert H.value(H.hsqr(x)) == pow(H.value(x), 2, p2)
            assert H.value(H.hadd(x, y)) == (H.value(x) + H.value(y)) % p2
        # extreme digits
        for x in [(H.R - 1, H.R - 1), (0, 0), (H.R - 1, 0), (0, H.R - 1), (p, p), (p - 1, 1)]:
            for y in [(H.R - 1, H.R - 1), (0, 0), (1, 0), (H.R - p, H.R - 1)]:
                assert H.value(H.hmul(x, y)) == (H.value(x) * H.value(y)) % p2
            assert H.value(H.hsqr(x)) == pow(H.value(x), 2, p2)
        # full decrypt side: c^(p-1) mod p^2, L function, times hp
        g = n + 1
        lg = (pow(g, p - 1, p2) - 1) // p
        hp = pow(lg, -1, p)
        for it in range(6 if Lh >= 16 else 40):
            msg = rnd.randrange(n)
            c = (pow(g, msg, n * n) * pow(rnd.randrange(1, n), n, n * n)) % (n * n)
            x = H.enter(c)
            assert H.value(x) == c % p2
            acc, nsq, nmul = hensel_pow(H, x, p - 1)
            want = pow(c, p - 1, p2)
            assert H.value(acc) == want
            mp, regular = H.finish(acc, hp)
            assert regular
            assert mp == (((want - 1) // p) * hp) % p == msg % p
        # a ciphertext that is a multiple of p is flagged irregular
        acc, _, _ = hensel_pow(H, H.enter(p * 12345), p - 1)
        assert not H.finish(acc, hp)[1]
        print("Lh=%d ok: max reduce rounds %d" % (Lh, H.max_rounds))
    print("schedule of a 1024-bit exponent: %d squarings + %d multiplies"
          % (nsq, nmul))
    print("model_hensel selftest ok")


if __name__ == "__main__":
    selftest()
