"""Word-level model of pailliercryptolib_b200/csrc/mont_hensel.cuh: the same
instruction sequences (32-bit registers, one carry flag per lane, T lanes in
lock step, shuffles and ballots) as HMont<K,T>::row / assemble / reduce /
pass_a / pass_b / sqr / mul / add, checked against the integer-level model in
tools/model_hensel.py.  Every register write asserts the 32-bit range, every
carry-out that the kernel drops is asserted to be zero."""
import random

from model_hensel import Hensel, gen_prime

M32 = (1 << 32) - 1


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def unlimbs(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


class Lane:
    """registers of one lane + its carry flag"""

    def __init__(self):
        self.cf = 0

    def add_cc(self, a, b):
        v = a + b
        self.cf = v >> 32
        return v & M32

    def addc_cc(self, a, b):
        v = a + b + self.cf
        self.cf = v >> 32
        return v & M32

    def addc(self, a, b):
        v = a + b + self.cf
        assert v <= M32, "addc dropped a carry"
        return v

    def addc_wrap(self, a, b):
        """addc whose carry-out is accounted elsewhere (ballot look-ahead)"""
        return (a + b + self.cf) & M32

    def mad_lo_cc(self, a, b, c):
        v = ((a * b) & M32) + c
        self.cf = v >> 32
        return v & M32

    def madc_lo_cc(self, a, b, c):
        v = ((a * b) & M32) + c + self.cf
        self.cf = v >> 32
        return v & M32

    def madc_hi_cc(self, a, b, c):
        v = ((a * b) >> 32) + c + self.cf
        self.cf = v >> 32
        return v & M32


class Group:
    def __init__(self, K, T):
        self.K, self.T = K, T
        self.lanes = [Lane() for _ in range(T)]

    def split(self, x):
        K, T = self.K, self.T
        l = limbs(x, K * T)
        return [l[t * K:(t + 1) * K] for t in range(T)]

    def join(self, v):
        return unlimbs([x for t in range(self.T) for x in v[t]])


def row(G, mode, P, Q, a, a2, n, b, b2, mi_l0, in_limb, n0inv):
    """P, Q, a, a2, n: per-lane lists.  Returns (new in_limb per lane, q)."""
    K, T = G.K, G.T
    for t in range(T):
        L = G.lanes[t]
        p, q = P[t], Q[t]
        t0 = L.add_cc(q[K], in_limb[t])
        t1 = L.addc(0, 0)
        p[0] = L.add_cc(p[0], q[1])
        for u in range(K // 2 - 1):
            q[2 * u] = L.madc_lo_cc(a[t][2 * u + 1], b, q[2 * u + 2])
            q[2 * u + 1] = L.madc_hi_cc(a[t][2 * u + 1], b, q[2 * u + 3])
        q[K - 2] = L.madc_lo_cc(a[t][K - 1], b, t0)
        q[K - 1] = L.madc_hi_cc(a[t][K - 1], b, t1)
        q[K] = L.addc(0, 0)
        p[0] = L.mad_lo_cc(a[t][0], b, p[0])
        p[1] = L.madc_hi_cc(a[t][0], b, p[1])
        for u in range(1, K // 2):
            p[2 * u] = L.madc_lo_cc(a[t][2 * u], b, p[2 * u])
            p[2 * u + 1] = L.madc_hi_cc(a[t][2 * u], b, p[2 * u + 1])
        p[K] = L.addc(p[K], 0)
        if mode == 2:
            p[0] = L.mad_lo_cc(a2[t][0], b2, p[0])
            p[1] = L.madc_hi_cc(a2[t][0], b2, p[1])
            for u in range(1, K // 2):
                p[2 * u] = L.madc_lo_cc(a2[t][2 * u], b2, p[2 * u])
                p[2 * u + 1] = L.madc_hi_cc(a2[t][2 * u], b2, p[2 * u + 1])
            p[K] = L.addc(p[K], 0)
            q[0] = L.mad_lo_cc(a2[t][1], b2, q[0])
            q[1] = L.madc_hi_cc(a2[t][1], b2, q[1])
            for u in range(1, K // 2):
                q[2 * u] = L.madc_lo_cc(a2[t][2 * u + 1], b2, q[2 * u])
                q[2 * u + 1] = L.madc_hi_cc(a2[t][2 * u + 1], b2, q[2 * u + 1])
            q[K] = L.addc(q[K], 0)
    mi0 = mi_l0 if mode else 0
    qd = ((P[0][0] + mi0) * n0inv) & M32  # lane 0, broadcast
    for t in range(T):
        L = G.lanes[t]
        p, q = P[t], Q[t]
        mi = mi0 if t == 0 else 0
        p[0] = L.mad_lo_cc(n[t][0], qd, p[0])
        p[1] = L.madc_hi_cc(n[t][0], qd, p[1])
        for u in range(1, K // 2):
            p[2 * u] = L.madc_lo_cc(n[t][2 * u], qd, p[2 * u])
            p[2 * u + 1] = L.madc_hi_cc(n[t][2 * u], qd, p[2 * u + 1])
        p[K] = L.addc(p[K], 0)
        if mode == 0:
            q[0] = L.mad_lo_cc(n[t][1], qd, q[0])
        else:
            L.add_cc(1 if mi else 0, M32)
            q[0] = L.madc_lo_cc(n[t][1], qd, q[0])
        q[1] = L.madc_hi_cc(n[t][1], qd, q[1])
        for u in range(1, K // 2):
            q[2 * u] = L.madc_lo_cc(n[t][2 * u + 1], qd, q[2 * u])
            q[2 * u + 1] = L.madc_hi_cc(n[t][2 * u + 1], qd, q[2 * u + 1])
        q[K] = L.addc(q[K], 0)
    # limb 0 of the group: 0 (pass A) or -mi (pass B), never read again
    assert (P[0][0] + mi0) & M32 == 0
    down = [P[t + 1][0] if t + 1 < T else 0 for t in range(T)]
    return down, qd


def resolve(G, r, g):
    K, T = G.K, G.T
    bg = bp = 0
    for t in range(T):
        if g[t]:
            bg |= 1 << t
        if all(x == M32 for x in r[t]):
            bp |= 1 << t
    cin = (bp + (bg << 1)) ^ bp
    for t in range(T):
        L = G.lanes[t]
        c = (cin >> t) & 1
        r[t][0] = L.add_cc(r[t][0], c)
        for j in range(1, K - 1):
            r[t][j] = L.addc_cc(r[t][j], 0)
        r[t][K - 1] = L.addc_wrap(r[t][K - 1], 0)
    return (cin >> T) & 1


def group_add(G, r, y, plus_one):
    K, T = G.K, G.T
    g = []
    for t in range(T):
        L = G.lanes[t]
        L.add_cc(plus_one if t == 0 else 0, M32)
        for j in range(K):
            r[t][j] = L.addc_cc(r[t][j], y[t][j])
        g.append(L.addc(0, 0))
    return resolve(G, r, g)


def assemble(G, E, O, t0, t1):
    K, T = G.K, G.T
    r, ov, g = [], [], []
    for t in range(T):
        L = G.lanes[t]
        rt = [0] * K
        rt[0] = L.add_cc(E[t][0], O[t][1])
        for j in range(1, K - 1):
            rt[j] = L.addc_cc(E[t][j], O[t][j + 1])
        rt[K - 1] = L.addc_cc(E[t][K - 1], t0[t])
        ov.append(L.addc(E[t][K], t1[t]))
        r.append(rt)
    for t in range(T):
        L = G.lanes[t]
        ov_in = ov[t - 1] if t else 0
        r[t][0] = L.add_cc(r[t][0], ov_in)
        for j in range(1, K):
            r[t][j] = L.addc_cc(r[t][j], 0)
        g.append(L.addc(0, 0))
    top = resolve(G, r, g)
    return r, ov[T - 1] + top


STATS = {"rounds": {}}


def reduce(G, r, c, dec, n):
    K, T = G.K, G.T
    subs = 0
    rounds = 0
    if c or dec:
        y = [[(~x) & M32 if c else (M32 if dec else 0) for x in n[t]] for t in range(T)]
        carry = group_add(G, r, y, 1 if (c and not dec) else 0)
        subs = 1 if c else 0
        c = c - 1 + carry
        assert c >= 0
        rounds += 1
    while c:
        y = [[(~x) & M32 for x in n[t]] for t in range(T)]
        carry = group_add(G, r, y, 1)
        subs += 1
        c = c - 1 + carry
        rounds += 1
        assert rounds <= 4
    STATS["rounds"][rounds] = STATS["rounds"].get(rounds, 0) + 1
    return subs


def pass_a(G, a, B, n, n0inv):
    K, T = G.K, G.T
    E = [[0] * (K + 1) for _ in range(T)]
    O = [[0] * (K + 1) for _ in range(T)]
    in_limb = [0] * T
    qs = []
    bl = limbs(B, K * T)
    for i in range(K * T):
        P, Q = (E, O) if i % 2 == 0 else (O, E)
        in_limb, qd = row(G, 0, P, Q, a, a, n, bl[i], 0, 0, in_limb, n0inv)
        qs.append(qd)
    t0, t1 = [], []
    for t in range(T):
        L = G.lanes[t]
        t0.append(L.add_cc(O[t][K], in_limb[t]))
        t1.append(L.addc(0, 0))
    r, c = assemble(G, E, O, t0, t1)
    assert c in (0, 1)
    reduce(G, r, c, 0, n)
    return r, c, qs


def pass_b(G, two, a, a2, B1, B0, qs, n, n0inv, hb, dec):
    K, T = G.K, G.T
    E = [[0] * (K + 1) for _ in range(T)]
    O = [[0] * (K + 1) for _ in range(T)]
    in_limb = [0] * T
    b1, b0 = limbs(B1, K * T), limbs(B0, K * T)
    for i in range(K * T):
        P, Q = (E, O) if i % 2 == 0 else (O, E)
        in_limb, _ = row(G, 2 if two else 1, P, Q, a, a2, n, b1[i], b0[i], qs[i],
                         in_limb, n0inv)
    t0, t1 = [], []
    for t in range(T):
        L = G.lanes[t]
        t0.append(L.add_cc(O[t][K], in_limb[t]))
        t1.append(L.addc(0, 0))
    if not two:
        for t in range(T):
            L = G.lanes[t]
            o, e = O[t], E[t]
            o[2] = L.mad_lo_cc(a[t][1], hb, o[2])
            o[3] = L.madc_hi_cc(a[t][1], hb, o[3])
            for u in range(1, K // 2 - 1):
                o[2 * u + 2] = L.madc_lo_cc(a[t][2 * u + 1], hb, o[2 * u + 2])
                o[2 * u + 3] = L.madc_hi_cc(a[t][2 * u + 1], hb, o[2 * u + 3])
            t0[t] = L.madc_lo_cc(a[t][K - 1], hb, t0[t])
            t1[t] = L.madc_hi_cc(a[t][K - 1], hb, t1[t])
            assert L.cf == 0
            e[0] = L.mad_lo_cc(a[t][0], hb, e[0])
            e[1] = L.madc_hi_cc(a[t][0], hb, e[1])
            for u in range(1, K // 2):
                e[2 * u] = L.madc_lo_cc(a[t][2 * u], hb, e[2 * u])
                e[2 * u + 1] = L.madc_hi_cc(a[t][2 * u], hb, e[2 * u + 1])
            e[K] = L.addc(e[K], 0)
    r, c = assemble(G, E, O, t0, t1)
    assert c <= 3
    reduce(G, r, c, dec, n)
    return r


def hsqr(G, x, p, n0inv):
    K, T = G.K, G.T
    R = 1 << (32 * K * T)
    x0, w = x
    a = G.split(x0)
    n = G.split(p)
    z0, ovA, qs = pass_a(G, a, x0, n, n0inv)
    dw, hb = (2 * w) % R, (2 * w) // R
    wz = pass_b(G, False, a, a, dw, dw, qs, n, n0inv, hb, ovA)
    return (G.join(z0), G.join(wz))


def hmul(G, x, y, p, n0inv):
    x0, w = x
    a, a2, n = G.split(x0), G.split(w), G.split(p)
    z0, ovA, qs = pass_a(G, a, y[0], n, n0inv)
    wz = pass_b(G, True, a, a2, y[1], y[0], qs, n, n0inv, 0, ovA)
    return (G.join(z0), G.join(wz))


def hadd(G, x, y, p):
    K, T = G.K, G.T
    n = G.split(p)
    x0, w = G.split(x[0]), G.split(x[1])
    c0 = group_add(G, x0, G.split(y[0]), 0)
    subs = reduce(G, x0, c0, 0, n)
    c = group_add(G, w, G.split(y[1]), 0)
    for it in range(3):
        if subs <= it:
            break
        yy = [list(n[t]) for t in range(T)]
        yy[0][0] -= 1
        c += group_add(G, w, yy, 0)
    reduce(G, w, c, 0, n)
    return (G.join(x0), G.join(w))


def selftest(seed=11):
    rnd = random.Random(seed)
    for (K, T) in [(4, 2), (8, 2), (16, 2), (24, 2), (12, 4), (16, 4)]:
        Lh = K * T
        R = 1 << (32 * Lh)
        iters = 60 if Lh <= 16 else 12
        for key in range(3):
            if key == 0:
                p = gen_prime(32 * Lh, rnd)
            elif key == 1:
                p = (R // 2) + 1 + 2 * rnd.randrange(1000)   # smallest allowed p
            else:
                p = R - 1 - 2 * rnd.randrange(1000)          # largest
            H = Hensel(p, Lh)
            n0inv = (-pow(p, -1, 1 << 32)) & M32
            G = Group(K, T)
            p2 = p * p
            for it in range(iters):
                mode = it % 5
                if mode == 0:
                    x = (R - 1, R - 1)
                    y = (R - 1, R - 1)
                elif mode == 1:
                    x = (rnd.randrange(R), 0)
                    y = (0, rnd.randrange(R))
                elif mode == 2:
                    x = (rnd.randrange(1 << 40), rnd.randrange(R))
                    y = (R - 1 - rnd.randrange(5), rnd.randrange(1 << 33))
                else:
                    x = (rnd.randrange(R), rnd.randrange(R))
                    y = (rnd.randrange(R), rnd.randrange(R))
                z = hmul(G, x, y, p, n0inv)
                assert z[0] < R and z[1] < R
                assert z == H.hmul(x, y), (K, T, key, it)
                assert H.value(z) == (H.value(x) * H.value(y)) % p2
                s = hsqr(G, x, p, n0inv)
                assert s == H.hsqr(x), (K, T, key, it)
                assert H.value(s) == pow(H.value(x), 2, p2)
                d = hadd(G, x, y, p)
                assert d[0] < R and d[1] < R
                assert H.value(d) == (H.value(x) + H.value(y)) % p2
        print("K=%d T=%d ok" % (K, T))
    tot = sum(STATS["rounds"].values())
    print("reduce rounds histogram:", {k: "%.1f%%" % (100.0 * v / tot)
                                      for k, v in sorted(STATS["rounds"].items())})
    print("model_hensel_words selftest ok")


if __name__ == "__main__":
    selftest()
