"""Bit-level Python model of the FP64-pipe Montgomery arithmetic in
pailliercryptolib_b200/csrc/mont_fp64.cuh (the DFMA half of the dual-pipe CRT
decrypt kernel).

A big integer is L = K*T limbs of W = 22 bits, each limb held as an
integer-valued double; lane t of a group of T lanes owns limbs [tK, (t+1)K).
Every product a_j*b_i is < 2^44.01 and is accumulated with fma() into a column
accumulator; a column receives at most 2 products per row for L rows, so with
L = 96 it stays below 2^52 and every fma is EXACT (no rounding ever happens;
the model asserts it).  Doubles are modelled as Python ints with that bound
asserted; the three places where the kernel uses floating-point rounding on
purpose (floor to a multiple of 2^W by adding 2^(52+W) in round-down mode, and
the 2^52 magic add that exposes the low mantissa bits) are modelled as floor.

Run:  python tools/model_fp64_mont.py
"""
import random

W = 22
MASK = (1 << W) - 1
LIM = 1 << 52


def chk(x):
    assert 0 <= x < LIM, "double accumulator out of the exact range: %d bits" % x.bit_length()
    return x


def to_limbs(x, L):
    out = [(x >> (W * i)) & MASK for i in range(L)]
    assert x >> (W * L) == 0
    return out


def from_limbs(v):
    return sum(x << (W * i) for i, x in enumerate(v))


def split(v, K, T):
    return [v[t * K:(t + 1) * K] for t in range(T)]


def join(lanes):
    return [x for ln in lanes for x in ln]


def normalize(lanes, K, T, passes=2):
    """lazy carry propagation: each pass moves floor(x / 2^W) to the next limb.
    After two passes every limb is < 2^W + 2^9."""
    for _ in range(passes):
        his = [[x >> W for x in ln] for ln in lanes]       # (x +rd 2^74) - 2^74
        los = [[x & MASK for x in ln] for ln in lanes]     # x - that
        for t in range(T):
            for j in range(K):
                if j > 0:
                    c = his[t][j - 1]
                elif t > 0:
                    c = his[t - 1][K - 1]                  # shuffle up
                else:
                    c = 0
                lanes[t][j] = chk(los[t][j] + c)           # fma(hi, 2^-W, lo)
        assert his[T - 1][K - 1] == 0, "carry out of the top limb"
    return lanes


def mont_mul(a, b, n, n0inv, K, T, rows=None):
    """a, n: per-lane limb lists; b: flat list of L limbs (the shared-memory
    operand).  Returns normalised per-lane limbs of a*b*2^(-W*rows) mod n."""
    L = K * T
    rows = L if rows is None else rows
    acc = [[0] * K for _ in range(T)]
    for i in range(rows):
        bi = b[i]
        for t in range(T):
            for j in range(K):
                acc[t][j] = chk(acc[t][j] + a[t][j] * bi)
        lo = chk(acc[0][0]) & 0xFFFFFFFF                    # low word of acc0 + 2^52
        q = (lo * n0inv) & MASK
        low = []
        for t in range(T):
            low.append(chk(acc[t][0] + n[t][0] * q))
            for j in range(1, K):
                acc[t][j - 1] = chk(acc[t][j] + n[t][j] * q)
        assert low[0] & MASK == 0
        for t in range(T):
            acc[t][K - 1] = low[t + 1] if t + 1 < T else 0  # shuffle down
        acc[0][0] = chk(acc[0][0] + (low[0] >> W))          # fma(low, 2^-W, acc0) in lane 0
    return normalize(acc, K, T)


def limbs_from_words(words, bitoff, count):
    """what the kernel does from the staged 32-bit words: `count` limbs of W
    bits starting at bit `bitoff`"""
    out = []
    for g in range(count):
        off = bitoff + W * g
        idx, sh = off >> 5, off & 31
        w0 = words[idx] if idx < len(words) else 0
        w1 = words[idx + 1] if idx + 1 < len(words) else 0
        out.append((((w1 << 32) | w0) >> sh) & MASK)
    return out


def pack_words(limbs, nwords):
    """exact carry propagation by one lane, then 32-bit words by all lanes"""
    limbs = list(limbs) + [0, 0, 0]
    c = 0
    for g in range(len(limbs)):
        v = limbs[g] + c
        limbs[g] = v & MASK
        c = v >> W
    assert c == 0
    words = []
    for k in range(nwords):
        g0 = (32 * k) // W
        o = 32 * k - W * g0
        v = limbs[g0] | (limbs[g0 + 1] << W) | ((limbs[g0 + 2] << (2 * W)) & ((1 << 64) - 1))
        words.append((v >> o) & 0xFFFFFFFF)
    return words


def build_schedule(e, w=5):
    """same format as build_schedule() in ipcl_b200.cu"""
    s = [1 << (w - 1)]
    bit = lambda i: i >= 0 and (e >> i) & 1
    i = e.bit_length() - 1
    first = True
    while i >= 0:
        if not bit(i):
            s.append(0)
            i -= 1
            continue
        l = max(i - w + 1, 0)
        while not bit(l):
            l += 1
        v = 0
        for k in range(i, l - 1, -1):
            v = (v << 1) | bit(k)
        if first:
            s.append((v - 1) // 2)
            first = False
        else:
            s += [0] * (i - l + 1)
            s.append((v - 1) // 2 + 1)
        i = l - 1
    s.append(0xFF)
    return s


def decrypt_side(ct, p, K=24, T=4, words32=64):
    """x = ct^(p-1) mod p^2 exactly as the FP64 role of the kernel computes it"""
    L = K * T
    n_int = p * p
    R = 1 << (W * L)
    assert 4 * n_int < R
    n = split(to_limbs(n_int, L), K, T)
    n0inv = (-pow(n_int, -1, 1 << W)) % (1 << W)
    r3 = to_limbs(pow(R, 3, n_int), L)
    ctw = [(ct >> (32 * i)) & 0xFFFFFFFF for i in range(2 * words32)]
    one = [1] + [0] * (L - 1)
    # prologue: ct = lo + hi*R;  ct*R^-1 = mont(lo, 1) + hi
    lo = split(limbs_from_words(ctw, 0, L), K, T)
    hi = split(limbs_from_words(ctw, W * L, L), K, T)
    x = mont_mul(lo, one, n, n0inv, K, T)
    x = [[chk(x[t][j] + hi[t][j]) for j in range(K)] for t in range(T)]
    x = normalize(x, K, T)
    x = mont_mul(x, r3, n, n0inv, K, T)                     # ct * R mod n
    assert from_limbs(join(x)) % n_int == ct * R % n_int
    sched = build_schedule(p - 1)
    nodd = sched[0]
    tab = [join(x)]
    x2 = join(mont_mul(x, join(x), n, n0inv, K, T))
    t = x
    for _ in range(1, nodd):
        t = mont_mul(t, x2, n, n0inv, K, T)
        tab.append(join(t))
    for e in tab:
        assert all(v < (1 << 23) for v in e)                # int32 table storage
    acc = split(tab[sched[1]], K, T)
    for op in sched[2:]:
        if op == 0xFF:
            break
        a = acc if op == 0 else split(tab[op - 1], K, T)
        acc = mont_mul(a, join(acc), n, n0inv, K, T)
        assert from_limbs(join(acc)) <= n_int + 1
    r = mont_mul(acc, one, n, n0inv, K, T)
    words = pack_words(join(r), words32)
    val = sum(w << (32 * i) for i, w in enumerate(words))
    assert val <= n_int
    if val == n_int:
        val = 0
    return val


def main():
    rnd = random.Random(22)
    K, T = 24, 4
    L = K * T
    # 1. single products, worst-case limbs
    for trial in range(20):
        n_int = rnd.getrandbits(2048) | (1 << 2047) | 1
        R = 1 << (W * L)
        n0inv = (-pow(n_int, -1, 1 << W)) % (1 << W)
        n = split(to_limbs(n_int, L), K, T)
        if trial == 0:
            a_int = b_int = 2 * n_int - 1
        else:
            a_int, b_int = rnd.randrange(2 * n_int), rnd.randrange(2 * n_int)
        a = split(to_limbs(a_int, L), K, T)
        r = mont_mul(a, to_limbs(b_int, L), n, n0inv, K, T)
        got = from_limbs(join(r))
        assert got % n_int == a_int * b_int * pow(R, -1, n_int) % n_int
        assert got <= n_int + 1
        assert all(v < (1 << W) + (1 << 9) for v in join(r))
    # worst-case lazy limbs everywhere (not a consistent number, bounds only)
    big = (1 << W) + (1 << 9) - 1
    n_int = (1 << 2048) - 1
    n = split(to_limbs(n_int, L), K, T)
    n0inv = (-pow(n_int, -1, 1 << W)) % (1 << W)
    try:
        mont_mul([[big] * K for _ in range(T)], [big] * L, n, n0inv, K, T)
    except AssertionError as e:
        # only the top-carry assertion may fire (the value is > 2n); the 2^52
        # accumulator bound must hold
        assert "carry out" in str(e) or not str(e), e
    print("mont_mul: ok (accumulators stay below 2^52)")
    # 2. a whole CRT side with a small 'prime' pair of the right size
    p = rnd.getrandbits(1024) | (1 << 1023) | 1
    for trial in range(2):
        ct = rnd.getrandbits(4096) % ((p * p) << 2040)
        assert decrypt_side(ct, p) == pow(ct, p - 1, p * p)
    print("decrypt side: ok")


if __name__ == "__main__":
    main()
