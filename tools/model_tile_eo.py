"""Python model of the even/odd-word window used by mont_tile.cuh (v2): the
8x8-limb tile is accumulated row-wise as carry chains of IMAD.WIDE on 64-bit
words (even words at even limb positions, odd words at odd positions), chain
carry-outs go to 32-bit counters.  Checks tile/double/resolve/shift against
plain integer arithmetic."""
import random

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1


class Win:
    def __init__(self):
        self.E = [0] * 8
        self.O = [0] * 8
        self.cE = [0] * 9
        self.cO = [0] * 9

    def value(self):
        v = 0
        for u in range(8):
            v += self.E[u] << (64 * u)
            v += self.O[u] << (64 * u + 32)
        for u in range(9):
            v += self.cE[u] << (64 * u)
            v += self.cO[u] << (64 * u + 32)
        return v

    def _chain(self, arr, cnt, base, x, ys):
        c = 0
        for t in range(4):
            s = arr[base + t] + x * ys[t] + c
            arr[base + t] = s & M64
            c = s >> 64
        cnt[base + 4] += c
        assert cnt[base + 4] <= M32

    def tile(self, X, Y):
        for i in range(8):
            ev, od = Y[0::2], Y[1::2]
            if i % 2 == 0:
                self._chain(self.E, self.cE, i // 2, X[i], ev)
                self._chain(self.O, self.cO, i // 2, X[i], od)
            else:
                self._chain(self.O, self.cO, i // 2, X[i], ev)
                self._chain(self.E, self.cE, (i + 1) // 2, X[i], od)

    def add_doubled(self, S):
        for arr, sarr, cnt, scnt in ((self.E, S.E, self.cE, S.cE), (self.O, S.O, self.cO, S.cO)):
            c = 0
            for u in range(8):
                d = ((sarr[u] << 1) | (sarr[u - 1] >> 63 if u else 0)) & M64
                s = arr[u] + d + c
                arr[u] = s & M64
                c = s >> 64
            cnt[8] += c + (sarr[7] >> 63)
            for u in range(9):
                cnt[u] += 2 * scnt[u]
                assert cnt[u] <= M32

    def add_block(self, P):
        # P: 8 limbs at limb 0..7, added into the E words (may carry)
        c = 0
        for u in range(4):
            s = self.E[u] + (P[2 * u] | (P[2 * u + 1] << 32)) + c
            self.E[u] = s & M64
            c = s >> 64
        self.cE[4] += c

    def resolve_low(self):
        T = []
        carry = 0
        for k in range(8):
            s = carry
            if k % 2 == 0:
                s += (self.E[k // 2] & M32) + self.cE[k // 2]
                if k:
                    s += self.O[k // 2 - 1] >> 32
            else:
                s += (self.E[k // 2] >> 32) + (self.O[k // 2] & M32) + self.cO[k // 2]
            T.append(s & M32)
            carry = s >> 32
        for u in range(4):
            self.E[u] = 0
            self.cE[u] = 0
            self.cO[u] = 0
        for u in range(3):
            self.O[u] = 0
        self.O[3] &= ~M32 & M64  # limb 8 part stays
        t = self.cE[4] + carry
        self.cE[4] = t & M32
        self.cO[4] += t >> 32
        assert self.cO[4] <= M32
        return T

    def set_low(self, T):
        for u in range(4):
            assert self.E[u] == 0
            self.E[u] = T[2 * u] | (T[2 * u + 1] << 32)

    def shift(self):
        """window moves up by 8 limbs; the low 8 limbs must be resolved (zero)"""
        x = self.O[3] >> 32  # limb 8 -> new limb 0
        self.E = self.E[4:] + [0] * 4
        self.O = self.O[4:] + [0] * 4
        self.cE = self.cE[4:] + [0] * 4
        self.cO = self.cO[4:] + [0] * 4
        t = self.cE[0] + x
        self.cE[0] = t & M32
        self.cO[0] += t >> 32


def selftest():
    rnd = random.Random(3)
    for it in range(300):
        W = Win()
        ref = 0
        S = Win()
        sref = 0
        for _ in range(rnd.randrange(1, 9)):
            X = [rnd.choice([M32, rnd.getrandbits(32)]) for _ in range(8)]
            Y = [rnd.choice([M32, rnd.getrandbits(32)]) for _ in range(8)]
            xv = sum(x << (32 * i) for i, x in enumerate(X))
            yv = sum(y << (32 * i) for i, y in enumerate(Y))
            if rnd.random() < 0.5:
                W.tile(X, Y)
                ref += xv * yv
            else:
                S.tile(X, Y)
                sref += xv * yv
        assert W.value() == ref and S.value() == sref
        W.add_doubled(S)
        ref += 2 * sref
        assert W.value() == ref
        P = [rnd.getrandbits(32) for _ in range(8)]
        W.add_block(P)
        ref += sum(x << (32 * i) for i, x in enumerate(P))
        assert W.value() == ref
        T = W.resolve_low()
        assert sum(t << (32 * i) for i, t in enumerate(T)) == ref & ((1 << 256) - 1)
        assert W.value() == ref - (ref & ((1 << 256) - 1))
        W.set_low(T)
        assert W.value() == ref
        W.resolve_low()
        W.shift()
        assert W.value() == ref >> 256
        # a second round on the shifted window
        X = [rnd.getrandbits(32) for _ in range(8)]
        W.tile(X, X)
        xv = sum(x << (32 * i) for i, x in enumerate(X))
        assert W.value() == (ref >> 256) + xv * xv
    print("model_tile_eo selftest ok")


if __name__ == "__main__":
    selftest()
