"""Bit-level Python model of the device Montgomery multiply in
pailliercryptolib_b200/csrc/mont_core.cuh (T lanes x K limbs, even/odd 64-bit
accumulator words, one-limb shift per row, lazy per-lane overflow words,
ballot carry look-ahead at the end).  Used to validate the algorithm and its
word-size bounds on the CPU before the kernel ever runs on a GPU."""
import random

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def unlimbs(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


class Lane:
    def __init__(self, K):
        self.K = K
        W = K // 2
        self.X = [0] * (W + 1)  # 64-bit words
        self.Y = [0] * (W + 1)


def row(lanes, pk, qk, a, n, b, in_limb, n0inv, K, T):
    """pk: name of even-role array, qk: odd-role array pending shift."""
    W = K // 2
    for t in range(T):
        L = lanes[t]
        P = getattr(L, pk)
        Q = getattr(L, qk)
        # top word of shifted Q: overflow word + incoming limb
        top = Q[W] + in_limb[t]
        assert top <= M64
        # fold hi32(Q[0]) into P[0].lo, carry into the odd chain
        lo = (P[0] & M32) + (Q[0] >> 32)
        c = lo >> 32
        P[0] = (P[0] & ~M32 & M64) | (lo & M32)
        # odd chain with right shift
        for u in range(W):
            src = Q[u + 1] if u < W - 1 else top
            v = a[t][2 * u + 1] * b + src + c
            c = v >> 64
            Q[u] = v & M64
        Q[W] = c
        # even chain
        c = 0
        for u in range(W):
            v = a[t][2 * u] * b + P[u] + c
            c = v >> 64
            P[u] = v & M64
        P[W] += c
        assert P[W] <= M64
    q = ((getattr(lanes[0], pk)[0] & M32) * n0inv) & M32
    out = []
    for t in range(T):
        L = lanes[t]
        P = getattr(L, pk)
        Q = getattr(L, qk)
        c = 0
        for u in range(W):
            v = n[t][2 * u] * q + P[u] + c
            c = v >> 64
            P[u] = v & M64
        P[W] += c
        c = 0
        for u in range(W):
            v = n[t][2 * u + 1] * q + Q[u] + c
            c = v >> 64
            Q[u] = v & M64
        Q[W] += c
        assert P[W] <= M64 and Q[W] <= M64
        out.append(P[0] & M32)
    assert out[0] == 0
    # limb moves one lane down
    return [out[t + 1] if t + 1 < T else 0 for t in range(T)]


def group_add(r, y, cin0, K, T):
    """r[t], y[t]: K-limb lists. returns (z, top_carry) with ballot look-ahead."""
    z = []
    g = 0
    p = 0
    for t in range(T):
        c = cin0 if t == 0 else 0
        zt = []
        for j in range(K):
            v = r[t][j] + y[t][j] + c
            zt.append(v & M32)
            c = v >> 32
        z.append(zt)
        if c:
            g |= 1 << t
        if all(x == M32 for x in zt):
            p |= 1 << t
    s = p + (g << 1)
    cin = (s ^ p)
    for t in range(T):
        c = (cin >> t) & 1
        for j in range(K):
            v = z[t][j] + c
            z[t][j] = v & M32
            c = v >> 32
    return z, (cin >> T) & 1


def montmul(A, B, N, n0inv, K, T):
    Lw = K * T
    W = K // 2
    a = [limbs(A, Lw)[t * K:(t + 1) * K] for t in range(T)]
    bl = limbs(B, Lw)
    n = [limbs(N, Lw)[t * K:(t + 1) * K] for t in range(T)]
    lanes = [Lane(K) for _ in range(T)]
    in_limb = [0] * T
    pk, qk = 'X', 'Y'
    for i in range(Lw):
        in_limb = row(lanes, pk, qk, a, n, bl[i], in_limb, n0inv, K, T)
        pk, qk = qk, pk
    # finalize: pk is the even-role (proper) array, qk pending shift
    r = []
    ovs = []
    for t in range(T):
        P = getattr(lanes[t], pk)
        Q = getattr(lanes[t], qk)
        e = []
        for u in range(W + 1):
            e += [P[u] & M32, P[u] >> 32]
        o = []
        for u in range(W + 1):
            o += [Q[u] & M32, Q[u] >> 32]
        # e[0..K+1] + o[1..K+1] at limb 0..K, + in_limb at K-1
        val = unlimbs(e) + unlimbs(o[1:K + 2]) + (in_limb[t] << (32 * (K - 1)))
        assert val < (1 << (32 * (K + 2)))
        rl = limbs(val, K + 2)
        r.append(rl[:K])
        ovs.append(rl[K] | (rl[K + 1] << 32))
    ovtop = ovs[T - 1]
    y = [[0] * K for _ in range(T)]
    for t in range(1, T):
        y[t][0] = ovs[t - 1] & M32
        y[t][1] = ovs[t - 1] >> 32
        assert ovs[t - 1] < (1 << 40)
    z, c1 = group_add(r, y, 0, K, T)
    ovf = ovtop + c1
    assert ovf in (0, 1), ovf
    if ovf:
        nn = [[(~x) & M32 for x in n[t]] for t in range(T)]
        z, _ = group_add(z, nn, 1, K, T)
    res = unlimbs([x for t in range(T) for x in z[t]])
    return res


def n0inv32(n):
    return (-pow(n, -1, 1 << 32)) & M32


def selftest(seed=1, iters=200):
    rnd = random.Random(seed)
    for (K, T) in [(2, 2), (4, 2), (8, 4), (16, 4), (16, 8), (12, 8), (24, 4)]:
        Lw = K * T
        R = 1 << (32 * Lw)
        for it in range(iters if K * T <= 64 else 20):
            mode = it % 6
            if mode == 0:
                N = R - 1 - 2 * rnd.randrange(1000)
            elif mode == 1:
                N = rnd.randrange(1, 1 << 40) | 1
            else:
                N = rnd.randrange(R >> rnd.randrange(1, 64), R) | 1
            if mode == 2:
                A, B = R - 1, R - 1
            elif mode == 3:
                A, B = R - 1 - rnd.randrange(5), rnd.randrange(R)
            else:
                A, B = rnd.randrange(R), rnd.randrange(R)
            got = montmul(A, B, N, n0inv32(N), K, T)
            assert got < R
            want = (A * B * pow(R, -1, N)) % N
            assert got % N == want, (K, T, it)
            # AMM bound: got == (A*B + q*N)/R possibly minus N
            full = (A * B + ((A * B * n0inv_full(N, R)) % R) * N) // R
            assert got in (full, full - N), (K, T, it)
    print("model_montmul selftest ok")


def n0inv_full(N, R):
    return (-pow(N, -1, R)) % R


if __name__ == "__main__":
    selftest()
