"""Python model of the thread-per-integer Montgomery multiply / square in
pailliercryptolib_b200/csrc/mont_tile.cuh: block-level finely integrated
product scanning with 8-limb blocks, 8x8 limb tiles accumulated column-wise
into 96-bit column accumulators, symmetric squaring (off-diagonal tiles once,
doubled), quotient block Q_c = low8(W * N') per column block."""
import random

M32 = (1 << 32) - 1
BL = 8  # limbs per block


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def unlimbs(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


class Cols:
    """15 column accumulators of 96 bits (lo, hi, ex kept as one int < 2^96)"""

    def __init__(self):
        self.c = [0] * 15

    def tile(self, X, Y):
        for i in range(BL):
            for j in range(BL):
                self.c[i + j] += X[i] * Y[j]
                assert self.c[i + j] < (1 << 96)

    def add_doubled(self, S):
        for k in range(15):
            self.c[k] += 2 * S.c[k]
            assert self.c[k] < (1 << 96)

    def add_block(self, P):
        for k in range(BL):
            self.c[k] += P[k]

    def resolve_low(self):
        """exact low 8 limbs; the carry moves into column 8"""
        out = []
        carry = 0
        for k in range(BL):
            t = self.c[k] + carry
            out.append(t & M32)
            carry = t >> 32
            self.c[k] = 0
        self.c[BL] += carry
        assert self.c[BL] < (1 << 96)
        return out

    def shift(self):
        self.c = self.c[BL:] + [0] * BL
        assert len(self.c) == 15


def low_mul(T, Ninv):
    """low 8 limbs of T * Ninv, column-wise"""
    out = []
    carry = 0
    for k in range(BL):
        s = carry
        for i in range(k + 1):
            s += T[i] * Ninv[k - i]
        out.append(s & M32)
        carry = s >> 32
    return out


def mont(A, B, N, Ninv8, NB, mode):
    """mode 'mul': A*B/R, 'sqr': A*A/R (symmetric), 'red': A is a 2*NB-block
    number, result A/R.  All almost-reduced (< R)."""
    L = NB * BL
    a = limbs(A, 2 * L if mode == 'red' else L)
    b = limbs(B, L) if mode == 'mul' else None
    n = limbs(N, L)
    blk = lambda v, i: v[i * BL:(i + 1) * BL]
    W = Cols()
    Q = [None] * NB
    R = []
    for c in range(2 * NB):
        if mode == 'sqr':
            S = Cols()
            for I in range(max(0, c - NB + 1), (c + 1) // 2):
                J = c - I
                assert I < J < NB
                S.tile(blk(a, I), blk(a, J))
            W.add_doubled(S)
            if c % 2 == 0 and c // 2 < NB:
                W.tile(blk(a, c // 2), blk(a, c // 2))
        elif mode == 'mul':
            for I in range(max(0, c - NB + 1), min(c, NB - 1) + 1):
                W.tile(blk(a, I), blk(b, c - I))
        else:
            W.add_block(blk(a, c))
        if c < NB:
            for I in range(0, c):
                W.tile(Q[I], blk(n, c - I))
            T = W.c[:]  # peek: resolve a copy of the low limbs
            low = []
            carry = 0
            for k in range(BL):
                t = T[k] + carry
                low.append(t & M32)
                carry = t >> 32
            Q[c] = low_mul(low, Ninv8)
            W.tile(Q[c], blk(n, 0))
            z = W.resolve_low()
            assert z == [0] * BL, z
        else:
            for I in range(c - NB + 1, NB):
                W.tile(Q[I], blk(n, c - I))
            R += W.resolve_low()
        W.shift()
    ovf = unlimbs([x & M32 for x in W.c[:1]]) + (W.c[0] >> 32)
    ovf = W.c[0]
    assert ovf in (0, 1), ovf
    r = unlimbs(R)
    if ovf:
        r = r + (1 << (32 * L)) - N
        assert r < (1 << (32 * L))
    return r


def selftest():
    rnd = random.Random(7)
    for NB in (2, 4, 8):
        L = NB * BL
        Rr = 1 << (32 * L)
        for it in range(60):
            N = (rnd.getrandbits(32 * L) | 1) if it % 3 else (Rr - 1 - 2 * rnd.randrange(99))
            if it % 5 == 4:
                N = rnd.getrandbits(40) | 1
            Ninv8 = limbs((-pow(N, -1, 1 << (32 * BL))) % (1 << (32 * BL)), BL)
            A = rnd.randrange(Rr) if it % 4 else Rr - 1
            B = rnd.randrange(Rr) if it % 4 else Rr - 1
            Rinv = pow(Rr, -1, N)
            got = mont(A, B, N, Ninv8, NB, 'mul')
            assert got < Rr and got % N == A * B * Rinv % N
            got = mont(A, None, N, Ninv8, NB, 'sqr')
            assert got < Rr and got % N == A * A * Rinv % N
            P = rnd.randrange(Rr * Rr)
            got = mont(P, None, N, Ninv8, NB, 'red')
            assert got < Rr and got % N == P * Rinv % N
    print("model_tile_fips selftest ok")


if __name__ == "__main__":
    selftest()
