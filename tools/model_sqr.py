"""Bit-level Python model of Mont<16,4>::sqr in
pailliercryptolib_b200/csrc/mont_sqr.cuh: Montgomery squaring of a 64-limb
integer spread over 4 lanes (16 limbs per lane) with the symmetric half of the
limb products.

  a = A0 + A1 X + A2 X^2 + A3 X^3  (X = 2^512);  a^2 needs the 4 diagonal
  blocks A_t^2 and the 6 off-diagonal blocks A_i A_j (i < j) doubled.
  Lane t computes   D_t  = A_t^2                      (in-lane, symmetric)
                    F_t  = 2 A_t A_{t+1 mod 4}        (16 rows)
                    Hf_t = 2 V rows8                   (8 rows: lanes 0,2 share
                           the block {0,2}, lanes 1,3 the block {1,3})
  i.e. 136 + 256 + 128 = 520 multiplies per lane instead of 1024; the partial
  products go through a per-group shared-memory area, every lane gathers the
  two 16-limb pieces of the 128-limb square it owns (W_t = limbs 16t.., H_t =
  limbs 64+16t..), and the Montgomery reduction runs on W alone (64 rows of
  q*n only); result = reduce(W) + H.

The model mirrors the kernel's data structures: even/odd 64-bit accumulator
words with carry counters for the static block products, 33-word slots, the
gather table, the two carry resolutions.  Run: python tools/model_sqr.py
"""
import random

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1
K, T = 16, 4
L = K * T


def limbs(x, n):
    out = [(x >> (32 * i)) & M32 for i in range(n)]
    assert x >> (32 * n) == 0
    return out


def val(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


def blockmul(V, rows):
    """static block product V (16 limbs) x rows (8 or 16 limbs) with even/odd
    64-bit accumulator words and carry counters; returns len(rows)+16 limbs
    (+1 for the final carry, always 0 here)"""
    nrows = len(rows)
    e = [0] * 34   # limb array of the even words: word m = limbs 2m, 2m+1 (positions 2m, 2m+1)
    o = [0] * 34   # limb array of the odd words:  word m = positions 2m+1, 2m+2
    ce = [0] * 18  # carry counters, position 2m
    co = [0] * 18  # position 2m+1

    def chain(arr, first_word, vs, b, cnt, ci):
        c = 0
        for s, v in enumerate(vs):
            m = first_word + s
            w = arr[2 * m] | (arr[2 * m + 1] << 32)
            w = w + v * b + c
            c = w >> 64
            w &= M64
            arr[2 * m], arr[2 * m + 1] = w & M32, w >> 32
        cnt[ci] += c

    for i in range(nrows):
        b = rows[i]
        r = i // 2
        if i % 2 == 0:
            chain(e, r, [V[2 * s] for s in range(8)], b, ce, r + 8)
            chain(o, r, [V[2 * s + 1] for s in range(8)], b, co, r + 8)
        else:
            chain(e, r + 1, [V[2 * s + 1] for s in range(8)], b, ce, r + 9)
            chain(o, r, [V[2 * s] for s in range(8)], b, co, r + 8)
    n_out = nrows + 16
    res = []
    c = 0
    for p in range(n_out + 1):
        t = e[p] + (o[p - 1] if p >= 1 else 0) + c
        res.append(t & M32)
        c = t >> 32
    assert c == 0
    c = 0
    for p in range(16, n_out + 1):
        cnt = ce[p // 2] if p % 2 == 0 else co[p // 2]
        t = res[p] + cnt + c
        res[p] = t & M32
        c = t >> 32
    assert c == 0
    assert val(res) == val(V) * val(rows)
    return res


def diag_square(V):
    """in-lane symmetric square: S = sum_{i<j} v_i v_j B^(i+j) with even/odd
    words and counters, doubled, plus the squares on the diagonal (one carry
    chain of 16 multiply-adds): 136 multiplies"""
    e = [0] * 34
    o = [0] * 34
    ce = [0] * 18
    co = [0] * 18
    for i in range(15):
        # j = i+1, i+3, ... : position i+j odd -> odd word (i+j-1)/2
        js = list(range(i + 1, 16, 2))
        c = 0
        for j in js:
            m = (i + j - 1) // 2
            w = (o[2 * m] | (o[2 * m + 1] << 32)) + V[i] * V[j] + c
            c = w >> 64
            w &= M64
            o[2 * m], o[2 * m + 1] = w & M32, w >> 32
        if js:
            co[(i + js[-1] - 1) // 2 + 1] += c
        # j = i+2, i+4, ... : position even -> even word (i+j)/2
        js = list(range(i + 2, 16, 2))
        c = 0
        for j in js:
            m = (i + j) // 2
            w = (e[2 * m] | (e[2 * m + 1] << 32)) + V[i] * V[j] + c
            c = w >> 64
            w &= M64
            e[2 * m], e[2 * m + 1] = w & M32, w >> 32
        if js:
            ce[(i + js[-1]) // 2 + 1] += c
    # S = e + (o << 32) + counters, 32 limbs
    s = []
    c = 0
    for p in range(32):
        t = e[p] + (o[p - 1] if p >= 1 else 0) + c
        s.append(t & M32)
        c = t >> 32
    assert c == 0
    c = 0
    for p in range(32):
        cnt = ce[p // 2] if p % 2 == 0 else co[p // 2]
        t = s[p] + cnt + c
        s[p] = t & M32
        c = t >> 32
    assert c == 0
    # double (funnel shift), then add v_i^2 at position 2i in one chain
    d = [((s[p] << 1) | (s[p - 1] >> 31 if p else 0)) & M32 for p in range(32)]
    assert s[31] >> 31 == 0
    c = 0
    for i in range(16):
        w = (d[2 * i] | (d[2 * i + 1] << 32)) + V[i] * V[i] + c
        c = w >> 64
        w &= M64
        d[2 * i], d[2 * i + 1] = w & M32, w >> 32
    assert c == 0
    assert val(d) == val(V) ** 2
    return d


def double(v, n_out):
    r = [((v[p] if p < len(v) else 0) << 1 | ((v[p - 1] >> 31) if 0 < p <= len(v) else 0)) & M32
         for p in range(n_out)]
    assert val(r) == 2 * val(v)
    return r


# gather table: half-block h of the 128-limb square <- (lane, slot kind, chunk)
# kinds: 0 = D, 1 = F, 2 = Hf;  chunk 0/1 = 16-limb pieces of the slot, 2 = the
# extra word (added at limb 0)
def gather_table():
    tab = {h: [] for h in range(8)}
    for t in range(4):
        # D_t at column 32 t
        tab[2 * t].append((t, 0, 0))
        tab[2 * t + 1].append((t, 0, 1))
        # F_t = 2 A_t A_{t+1}: column 16 (t + (t+1)%4)
        hb = t + (t + 1) % 4
        tab[hb].append((t, 1, 0))
        tab[hb + 1].append((t, 1, 1))
        if hb + 2 < 8:
            tab[hb + 2].append((t, 1, 2))
        # Hf_t: lanes 0,2 -> block {0,2} (column 32), lanes 1,3 -> {1,3} (column 64);
        # the slot is positioned at the block's column (result stored at offset 0 or 8)
        hb = 2 if t % 2 == 0 else 4
        tab[hb].append((t, 2, 0))
        tab[hb + 1].append((t, 2, 1))
        tab[hb + 2].append((t, 2, 2))
    return tab


def sqr_product(a_int):
    """phase 1 + gather: returns per-lane W, H (16 limbs each) of a^2"""
    a = limbs(a_int, L)
    A = [a[16 * t:16 * t + 16] for t in range(4)]
    slots = {}
    for t in range(4):
        d = diag_square(A[t])
        slots[(t, 0)] = d + [0]
        f = double(blockmul(A[t], A[(t + 1) % 4]), 33)
        slots[(t, 1)] = f
        V = A[0] if t % 2 == 0 else A[1]
        src = A[2] if t % 2 == 0 else A[3]
        i0 = 0 if t < 2 else 8
        h = double(blockmul(V, src[i0:i0 + 8]), 25)
        s = [0] * 33
        for k, x in enumerate(h):
            s[i0 + k] = x
        slots[(t, 2)] = s
    tab = gather_table()
    maxc = max(len(v) for v in tab.values())
    piece = {}
    cnt = {}
    for h in range(8):
        acc = [0] * 16
        c_out = 0
        x0 = 0
        for (t, kind, ch) in tab[h]:
            s = slots[(t, kind)]
            if ch == 2:
                x0 += s[32]
                continue
            c = 0
            for j in range(16):
                v = acc[j] + s[16 * ch + j] + c
                acc[j] = v & M32
                c = v >> 32
            c_out += c
        c = x0
        for j in range(16):
            v = acc[j] + c
            acc[j] = v & M32
            c = v >> 32
        c_out += c
        piece[h] = acc
        cnt[h] = c_out
    # carries between the pieces: cnt[h] enters piece h+1 at limb 0
    carry = 0
    full = []
    for h in range(8):
        c = carry
        for j in range(16):
            v = piece[h][j] + c
            piece[h][j] = v & M32
            c = v >> 32
        carry = cnt[h] + c
        full += piece[h]
    assert carry == 0
    assert val(full) == a_int * a_int
    return full[:64], full[64:], maxc


def mont_sqr(a_int, n_int):
    R = 1 << (32 * L)
    W, H, _ = sqr_product(a_int)
    w, h = val(W), val(H)
    npr = (-pow(n_int, -1, R)) % R
    q = (w * npr) % R
    v = (w + q * n_int) // R          # 64 reduce-only rows
    r = v + h
    if r >= R:
        r -= n_int
    assert r < R
    assert (r * R - a_int * a_int) % n_int == 0
    return r


# ---------------------------------------------------------------------------
# the 32 x 2 layout (MontSqr2): lane t holds A_t (32 limbs), a = A0 + A1 X
#   D_t = A_t^2 (in-lane symmetric, 528 multiplies)
#   C_t = 2 A0 * A1[16t .. 16t+16)  (512 multiplies; lane 1 reads A0 from shared memory)
# ---------------------------------------------------------------------------
def blockmul_g(V, rows):
    """generic static block product: len(V) even, any number of rows"""
    kv = len(V) // 2
    n_out = len(rows) + len(V)
    e = [0] * (n_out + 4)
    o = [0] * (n_out + 4)
    ce = [0] * (n_out // 2 + 4)
    co = [0] * (n_out // 2 + 4)

    def chain(arr, first_word, vs, b, cnt, ci):
        c = 0
        for s_, v in enumerate(vs):
            m = first_word + s_
            w = (arr[2 * m] | (arr[2 * m + 1] << 32)) + v * b + c
            c = w >> 64
            w &= M64
            arr[2 * m], arr[2 * m + 1] = w & M32, w >> 32
        cnt[ci] += c

    for i, b in enumerate(rows):
        r = i // 2
        if i % 2 == 0:
            chain(e, r, [V[2 * s_] for s_ in range(kv)], b, ce, r + kv)
            chain(o, r, [V[2 * s_ + 1] for s_ in range(kv)], b, co, r + kv)
        else:
            chain(e, r + 1, [V[2 * s_ + 1] for s_ in range(kv)], b, ce, r + kv + 1)
            chain(o, r, [V[2 * s_] for s_ in range(kv)], b, co, r + kv)
    res = []
    c = 0
    for p in range(n_out + 1):
        t = e[p] + (o[p - 1] if p >= 1 else 0) + c
        res.append(t & M32)
        c = t >> 32
    assert c == 0
    c = 0
    for p in range(n_out + 1):
        cnt = ce[p // 2] if p % 2 == 0 else co[p // 2]
        t = res[p] + cnt + c
        res[p] = t & M32
        c = t >> 32
    assert c == 0
    assert val(res) == val(V) * val(rows)
    return res


def diag_square_g(V):
    n = len(V)
    e = [0] * (2 * n + 4)
    o = [0] * (2 * n + 4)
    ce = [0] * (n + 4)
    co = [0] * (n + 4)
    for i in range(n - 1):
        for start, arr, cnt, odd in ((i + 1, o, co, 1), (i + 2, e, ce, 0)):
            js = list(range(start, n, 2))
            c = 0
            m = None
            for j in js:
                m = (i + j - odd) // 2
                w = (arr[2 * m] | (arr[2 * m + 1] << 32)) + V[i] * V[j] + c
                c = w >> 64
                w &= M64
                arr[2 * m], arr[2 * m + 1] = w & M32, w >> 32
            if js:
                cnt[m + 1] += c
    s = []
    c = 0
    for p in range(2 * n):
        t = e[p] + (o[p - 1] if p >= 1 else 0) + c
        s.append(t & M32)
        c = t >> 32
    assert c == 0
    c = 0
    for p in range(2 * n):
        cnt = ce[p // 2] if p % 2 == 0 else co[p // 2]
        t = s[p] + cnt + c
        s[p] = t & M32
        c = t >> 32
    assert c == 0
    d = [((s[p] << 1) | (s[p - 1] >> 31 if p else 0)) & M32 for p in range(2 * n)]
    c = 0
    for i in range(n):
        w = (d[2 * i] | (d[2 * i + 1] << 32)) + V[i] * V[i] + c
        c = w >> 64
        w &= M64
        d[2 * i], d[2 * i + 1] = w & M32, w >> 32
    assert c == 0 and val(d) == val(V) ** 2
    return d


# shared-memory words of one group in the 32 x 2 layout: a (64), then per lane
# a D slot (64 limbs) and a C slot (64 limbs + extra word + pad = 68)
SQ2_D = lambda t: 64 + 132 * t
SQ2_C = lambda t: 64 + 132 * t + 64
SQ2_CHUNK = {0: [SQ2_D(0)], 1: [SQ2_D(0) + 32, SQ2_C(0), SQ2_C(1)],
             2: [SQ2_D(1), SQ2_C(0) + 32, SQ2_C(1) + 32], 3: [SQ2_D(1) + 32]}
SQ2_EXTRA = {0: [], 1: [], 2: [], 3: [SQ2_C(1) + 64]}


def sqr2_product(a_int):
    a = limbs(a_int, 64)
    A = [a[:32], a[32:]]
    mem = [0] * 332
    for t in range(2):
        d = diag_square_g(A[t])
        mem[SQ2_D(t):SQ2_D(t) + 64] = d
        rows = A[1][16 * t:16 * t + 16]
        c = double(blockmul_g(A[0], rows), 49)
        off = SQ2_C(t) + 16 * t
        mem[off:off + 49] = c
    piece, cnt = {}, {}
    for h in range(4):
        acc = [0] * 32
        c_out = 0
        for off in SQ2_CHUNK[h]:
            c = 0
            for j in range(32):
                v = acc[j] + mem[off + j] + c
                acc[j] = v & M32
                c = v >> 32
            c_out += c
        c = sum(mem[off] for off in SQ2_EXTRA[h])
        for j in range(32):
            v = acc[j] + c
            acc[j] = v & M32
            c = v >> 32
        c_out += c
        piece[h], cnt[h] = acc, c_out
    carry = 0
    full = []
    for h in range(4):
        c = carry
        for j in range(32):
            v = piece[h][j] + c
            piece[h][j] = v & M32
            c = v >> 32
        carry = cnt[h] + c
        full += piece[h]
    assert carry == 0 and val(full) == a_int * a_int
    return full


def main():
    rnd = random.Random(16)
    for a in [(1 << 2048) - 1, 0, int("ffffffff00000000" * 32, 16)] + \
            [rnd.getrandbits(2048) for _ in range(12)]:
        sqr2_product(a)
    print("model_sqr (32 x 2 layout): ok")
    for trial in range(30):
        if trial == 0:
            a = (1 << 2048) - 1
        elif trial == 1:
            a = 0
        elif trial == 2:
            a = int("ffffffff00000000" * 32, 16)
        else:
            a = rnd.getrandbits(2048)
        W, H, maxc = sqr_product(a)
        n = rnd.getrandbits(2048) | (1 << 2047) | 1
        mont_sqr(a, n)
    tab = gather_table()
    print("gather table (half-block <- lane, kind, chunk):")
    for h in range(8):
        print("  ", h, tab[h])
    print("model_sqr: ok, max contributions per half-block:", maxc)


if __name__ == "__main__":
    main()
