"""Parity of the CUDA path (through the C ABI) against the oracle, the golden
fixtures and the reference's ISO/IEC 18033-6 known-answer vectors.  Bit-exact:
integer work, no tolerance."""
import math
import os

import numpy as np
import pytest

from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,
                                          from_limbs, random_limbs, to_limbs)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default-layout", "wide-layout", "mid-layout"])
def layout(request, monkeypatch):
    """runs a test with the default lane layout (16 or 12 limbs per lane), the
    small-batch layout (4 limbs per lane, four times as many lanes per integer)
    and the one in between (8 limbs per lane); without the override the library
    picks by batch size"""
    monkeypatch.setenv("IPCLB200_WIDE", {"default-layout": "0", "wide-layout": "1",
                                         "mid-layout": "2"}[request.param])
    return request.param


def test_modexp_golden_vectors(capi, modexp_vectors, layout):
    """every golden vector, one modulus per call group (heterogeneous path)"""
    by_bits = {}
    for v in modexp_vectors:
        by_bits.setdefault(v["bits"], []).append(v)
    for bits, vs in by_bits.items():
        L = bits // 32
        b = batch_to_limbs([int(v["b"], 16) for v in vs], L)
        e = batch_to_limbs([int(v["e"], 16) for v in vs], L)
        m = batch_to_limbs([int(v["m"], 16) for v in vs], L)
        r = capi.modexp(b, e, m, 0)
        assert batch_from_limbs(r) == [int(v["r"], 16) for v in vs], bits


@pytest.mark.parametrize("bits", [512, 1024, 1536, 2048, 3072, 4096, 6144, 8192])
def test_modexp_random_vs_oracle(capi, oracle, bits, layout):
    """shared odd modulus with the top bit set, per-element base and exponent;
    batch sizes that are not multiples of the group count"""
    L = bits // 32
    rng = np.random.default_rng(bits)
    count = {512: 1500, 1024: 700, 1536: 300, 2048: 301, 3072: 100, 4096: 67,
             6144: 21, 8192: 9}[bits]
    mod = random_limbs(rng, 1, L)
    mod[0, 0] |= 1
    mod[0, -1] |= 0x80000000
    base = random_limbs(rng, count, L)          # includes bases >= modulus
    ebits_words = max(1, L // 2)
    exp = random_limbs(rng, count, ebits_words)
    got = capi.modexp(base, exp, mod, capi.SHARED_MOD)
    want = oracle.modexp(base, exp, mod, shared_mod=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("words", [1, 7, 16, 20, 33, 40, 65, 100, 129])
def test_modexp_unaligned_widths(capi, oracle, words):
    """moduli whose width is not a kernel size class are zero padded"""
    rng = np.random.default_rng(words)
    count = 37
    mod = random_limbs(rng, 1, words, top_mask=0x0FFFFFFF)
    mod[0, 0] |= 1
    mod[0, -1] |= 0x01000000
    base = random_limbs(rng, count, words)
    exp = random_limbs(rng, count, 3)
    got = capi.modexp(base, exp, mod, capi.SHARED_MOD)
    want = oracle.modexp(base, exp, mod, shared_mod=True)
    assert np.array_equal(got, want)


def test_modexp_shared_operands(capi, oracle):
    rng = np.random.default_rng(7)
    L, count = 64, 50
    mod = random_limbs(rng, 1, L)
    mod[0, 0] |= 1
    base1 = random_limbs(rng, 1, L)
    exp1 = random_limbs(rng, 1, 8)
    base = random_limbs(rng, count, L)
    exp = random_limbs(rng, count, 8)
    got = capi.modexp(base1, exp, mod, capi.SHARED_MOD | capi.SHARED_BASE)
    want = oracle.modexp(base1, exp, mod, shared_mod=True, shared_base=True)
    assert np.array_equal(got, want)
    got = capi.modexp(base, exp1, mod, capi.SHARED_MOD | capi.SHARED_EXP)
    want = oracle.modexp(base, exp1, mod, shared_mod=True, shared_exp=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ebits", [65, 700, 2048])
def test_modexp_shared_exponent_schedule(capi, ebits, monkeypatch):
    """one exponent for the whole batch (IPCLB200_SHARED_EXP, count >= 64) runs
    the host-built sliding-window schedule; same residues as the fixed-window
    scan and as Python pow(), including exponents 2^k and 2^k - 1"""
    rng = np.random.default_rng(ebits)
    L, count = 64, 100
    mod = random_limbs(rng, 1, L)
    mod[0, 0] |= 1
    mod[0, -1] |= 0x80000000
    m = from_limbs(mod[0])
    base = random_limbs(rng, count, L)
    B = batch_from_limbs(base)
    ew = (ebits + 31) // 32
    for e in (int.from_bytes(rng.bytes(ew * 4), "little") >> (32 * ew - ebits) | (1 << (ebits - 1)),
              1 << (ebits - 1), (1 << ebits) - 1):
        el = to_limbs(e, ew)[None, :]
        got = capi.modexp(base, el, mod, capi.SHARED_MOD | capi.SHARED_EXP)
        assert batch_from_limbs(got) == [pow(b, e, m) for b in B]
        monkeypatch.setenv("IPCLB200_NO_SCHED", "1")
        assert np.array_equal(capi.modexp(base, el, mod, capi.SHARED_MOD | capi.SHARED_EXP), got)
        monkeypatch.delenv("IPCLB200_NO_SCHED")


def test_modexp_edge_exponents_and_batch_one(capi):
    L = 32
    m = (1 << 1023) + 1155
    mod = batch_to_limbs([m], L)
    for b, e in [(5, 0), (0, 0), (0, 5), (1, 12345), (m - 1, 2), (m - 1, 3),
                 (m + 7, 3), (2, 1 << 100)]:
        got = capi.modexp(batch_to_limbs([b], L), batch_to_limbs([e], 4), mod,
                          capi.SHARED_MOD)
        assert from_limbs(got[0]) == pow(b, e, m), (b, e)
    one = batch_to_limbs([1], L)
    got = capi.modexp(batch_to_limbs([5], L), batch_to_limbs([3], 1), one,
                      capi.SHARED_MOD)
    assert from_limbs(got[0]) == 0


def test_modexp_errors(capi):
    L = 16
    b = batch_to_limbs([3], L)
    with pytest.raises(capi.IpclB200Error) as ei:
        capi.modexp(b, b, batch_to_limbs([10], L), capi.SHARED_MOD)
    assert ei.value.code == -2
    with pytest.raises(capi.IpclB200Error) as ei:
        capi.modexp(b, b, batch_to_limbs([0], L), capi.SHARED_MOD)
    assert ei.value.code == -1
    # empty batch is a no-op
    out = capi.modexp(np.zeros((0, L), np.uint32), np.zeros((0, L), np.uint32),
                      batch_to_limbs([7], L), capi.SHARED_MOD)
    assert out.shape[0] <= 1


@pytest.mark.parametrize("bits", [1024, 2048, 4096, 6144])
def test_modmul_vs_oracle(capi, oracle, bits):
    L = bits // 32
    rng = np.random.default_rng(bits + 1)
    count = 333
    for top in (0x80000000, 0x00000001):       # large and "small" modulus
        mod = random_limbs(rng, 1, L, top_mask=(top << 1) - 1 if top > 1 else 1)
        mod[0, 0] |= 1
        mod[0, -1] |= top
        a = random_limbs(rng, count, L)
        b = random_limbs(rng, count, L)
        got = capi.modmul(a, b, mod[0])
        want = oracle.modmul(a, b, mod[0])
        assert np.array_equal(got, want)
        got = capi.modmul(a, b[0:1], mod[0], capi.SHARED_B)
        want = oracle.modmul(a, b[0:1], mod[0], b_shared=True)
        assert np.array_equal(got, want)


def test_iso_kat_through_cabi(capi, iso, layout):
    """The reference's only known-answer test, replayed through the C ABI:
    test_cryptography.cpp:198-240 (21 values, non-DJN key, injected r)."""
    p, q = iso["p"], iso["q"]
    n = p * q
    pk = capi.PubKey(to_limbs(n, 64))
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    ms = [iso["m0"]] * 21
    rs = [iso["r0"]] * 21
    ms[1], rs[1] = iso["m1"], iso["r1"]
    ct = pk.encrypt(batch_to_limbs(ms, 64), batch_to_limbs(rs, 64))
    got = batch_from_limbs(ct)
    assert got[0] == iso["c1"]
    assert got[1] == iso["c2"]
    assert batch_from_limbs(sk.decrypt(ct)) == ms
    s = capi.modmul(ct[0:1], ct[1:2], to_limbs(n * n, 128))
    assert from_limbs(s[0]) == iso["c1c2"]
    assert from_limbs(sk.decrypt(s)[0]) == iso["m1m2"]
    assert from_limbs(sk.decrypt(s, use_crt=False)[0]) == iso["m1m2"]


@pytest.mark.parametrize("bits", ["1024", "2048", "3072"])
def test_scheme_golden_through_cabi(capi, keys, scheme_vectors, bits, monkeypatch):
    k = keys[bits]
    items = scheme_vectors[bits]
    p, q, hs = k["p"], k["q"], k["hs"]
    n = p * q
    NL = int(bits) // 32
    ms = [int(i["m"], 16) for i in items]
    pt = batch_to_limbs(ms, NL)
    r_djn = batch_to_limbs([int(i["r_djn"], 16) for i in items], NL // 2)
    r_std = batch_to_limbs([int(i["r_std"], 16) for i in items], NL)
    pk_djn = capi.PubKey(to_limbs(n, NL), to_limbs(hs, 2 * NL), int(bits) // 2)
    pk_std = capi.PubKey(to_limbs(n, NL))
    sk = capi.PrivKey(to_limbs(q, NL // 2), to_limbs(p, NL // 2))  # swapped on purpose
    c_djn = pk_djn.encrypt(pt, r_djn)         # fixed-base comb for hs^r
    monkeypatch.setenv("IPCLB200_NO_COMB", "1")
    c_djn_win = pk_djn.encrypt(pt, r_djn)     # generic fixed-window hs^r
    monkeypatch.delenv("IPCLB200_NO_COMB")
    assert np.array_equal(c_djn, c_djn_win)
    c_std = pk_std.encrypt(pt, r_std)
    c_plain = pk_std.encrypt(pt, None, make_secure=False)
    assert batch_from_limbs(c_djn) == [int(i["c_djn"], 16) for i in items]
    assert batch_from_limbs(c_std) == [int(i["c_std"], 16) for i in items]
    assert batch_from_limbs(c_plain) == [int(i["c_plain"], 16) for i in items]
    assert batch_from_limbs(sk.decrypt(c_djn)) == ms
    assert batch_from_limbs(sk.decrypt(c_std, use_crt=False)) == ms
    nsq = to_limbs(n * n, 2 * NL)
    assert batch_from_limbs(capi.modmul(c_djn, c_std, nsq)) == \
        [int(i["c_add"], 16) for i in items]
    kk = batch_to_limbs([int(i["k"], 16) for i in items], 2 * NL)
    assert batch_from_limbs(capi.modexp(c_djn, kk, nsq, capi.SHARED_MOD)) == \
        [int(i["c_mul"], 16) for i in items]


@pytest.mark.parametrize("bits", ["1024", "2048", "3072"])
def test_encrypt_decrypt_batch_vs_oracle(capi, oracle, keys, bits, layout):
    """larger batch: DJN through the fixed-base comb (count >= 64), non-DJN,
    decrypt CRT and RAW, all against the oracle on the same seeded inputs"""
    k = keys[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    count = {"1024": 531, "2048": 203, "3072": 77}[bits]
    rng = np.random.default_rng(int(bits))
    pt = batch_to_limbs([int.from_bytes(rng.bytes(NL * 4), "little") % n
                         for _ in range(count)], NL)
    r_djn = random_limbs(rng, count, NL // 2)
    r_std = batch_to_limbs([1 + int.from_bytes(rng.bytes(NL * 4), "little") % (n - 1)
                            for _ in range(count)], NL)
    nl = to_limbs(n, NL)
    hsl = to_limbs(k["hs"], 2 * NL)
    pk_djn = capi.PubKey(nl, hsl, int(bits) // 2)
    pk_std = capi.PubKey(nl)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    c_djn = pk_djn.encrypt(pt, r_djn)
    assert np.array_equal(c_djn, oracle.encrypt(nl, hsl, pt, r_djn))
    c_std = pk_std.encrypt(pt[:40], r_std[:40])
    assert np.array_equal(c_std, oracle.encrypt(nl, None, pt[:40], r_std[:40]))
    d = sk.decrypt(c_djn)
    assert np.array_equal(d, pt)
    assert np.array_equal(d, oracle.decrypt_crt(to_limbs(p, NL // 2),
                                                to_limbs(q, NL // 2), c_djn))
    d_raw = sk.decrypt(c_djn[:40], use_crt=False)
    assert np.array_equal(d_raw, pt[:40])
    # wide injected randoms (the reference benchmark injects a 2047-bit r with
    # DJN on, bench_cryptography.cpp:38-47,81-82): comb table is extended
    r_wide = random_limbs(rng, 70, NL)
    c_w = pk_djn.encrypt(pt[:70], r_wide)
    assert np.array_equal(c_w, oracle.encrypt(nl, hsl, pt[:70], r_wide))


@pytest.mark.parametrize("bits,count", [("1024", 2500), ("2048", 2051)])
def test_non_djn_encrypt_two_digit_ladder_vs_oracle(capi, oracle, keys, bits, count, monkeypatch):
    """non-DJN keys (the reference's default PublicKey): obf = r^n mod n^2
    (ipcl/pub_key.cpp:66-80).  From 2048 elements on the library computes it with
    the two-digit ladder + n*m+1 + one modmul instead of the full-width fused
    kernel: same ciphertexts from both, equal to the oracle, decryptable"""
    k = keys[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    rng = np.random.default_rng(int(bits) + count)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = batch_to_limbs([1 + int.from_bytes(rng.bytes(NL * 4), "little") % (n - 1)
                        for _ in range(count)], NL)
    nl = to_limbs(n, NL)
    pk = capi.PubKey(nl)
    ct = pk.encrypt(pt, r)
    monkeypatch.setenv("IPCLB200_NO_HENSEL_MODEXP", "1")
    assert np.array_equal(pk.encrypt(pt, r), ct)  # the full-width fused kernel
    monkeypatch.delenv("IPCLB200_NO_HENSEL_MODEXP")
    assert np.array_equal(ct[:300], oracle.encrypt(nl, None, pt[:300], r[:300]))
    assert np.array_equal(ct[-5:], oracle.encrypt(nl, None, pt[-5:], r[-5:]))
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    assert np.array_equal(sk.decrypt(ct), pt)
    # RAW decrypt (ipcl/pri_key.cpp:92-111) takes the same ladder for ct^lambda
    assert np.array_equal(sk.decrypt(ct, use_crt=False), pt)
    monkeypatch.setenv("IPCLB200_NO_HENSEL_MODEXP", "1")
    assert np.array_equal(sk.decrypt(ct[:2048], use_crt=False), pt[:2048])


def test_comb_table_upgrade_keeps_results(capi, oracle, keys, monkeypatch):
    """a DJN key starts with the small fixed-base table (8-bit windows) and moves
    to the wide one after IPCLB200_COMB_UPGRADE elements: same ciphertexts from
    both, equal to the oracle"""
    k = keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    rng = np.random.default_rng(808)
    pt = random_limbs(rng, 300, 64, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, 300, 32)
    nl, hsl = to_limbs(n, 64), to_limbs(k["hs"], 128)
    monkeypatch.setenv("IPCLB200_COMB_UPGRADE", "500")
    pk = capi.PubKey(nl, hsl, 1024)
    c_small = pk.encrypt(pt, r)     # 300 elements so far: starter table
    c_wide = pk.encrypt(pt, r)      # 600: upgraded
    c_again = pk.encrypt(pt, r)
    want = oracle.encrypt(nl, hsl, pt, r)
    assert np.array_equal(c_small, want)
    assert np.array_equal(c_wide, want)
    assert np.array_equal(c_again, want)


def test_homomorphic_properties_large_batch(capi, keys):
    """size-independent checks at a batch the oracle would need minutes for:
    dec(enc(a) * enc(b)) = a + b mod n, dec(enc(a)^k) = a*k mod n."""
    k = keys["2048"]
    p, q = k["p"], k["q"]
    n = p * q
    NL = 64
    count = 4096
    rng = np.random.default_rng(2048)
    a = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    b = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r1 = random_limbs(rng, count, 32)
    r2 = random_limbs(rng, count, 32)
    kk = random_limbs(rng, count, 1)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    ca, cb = pk.encrypt(a, r1), pk.encrypt(b, r2)
    nsq = to_limbs(n * n, 128)
    s = sk.decrypt(capi.modmul(ca, cb, nsq))
    m = sk.decrypt(capi.modexp(ca, kk, nsq, capi.SHARED_MOD))
    A, B, S, M = (batch_from_limbs(x) for x in (a, b, s, m))
    K = [int(x) for x in kk[:, 0]]
    assert S == [(x + y) % n for x, y in zip(A, B)]
    assert M == [(x * y) % n for x, y in zip(A, K)]
    assert np.array_equal(sk.decrypt(ca), a)


def test_int_peak_reports(capi):
    macs, mhz = capi.int_peak()
    assert macs > 1e12 and mhz > 500


needs_experiments = pytest.mark.skipif(
    os.environ.get("IPCLB200_EXPERIMENTS", "0") != "1",
    reason="libipcl_b200.so is built without -DIPCLB200_EXPERIMENTS (build.py --experiments)")


@needs_experiments
@pytest.mark.parametrize("bits", ["1024", "2048"])
def test_decrypt_tile_kernel_matches(capi, keys, bits, monkeypatch):
    """the opt-in thread-per-integer decrypt kernel (mont_tile.cuh) gives the
    same plaintexts as the default lane-distributed one"""
    k = keys[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    rng = np.random.default_rng(99)
    count = 333
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, NL // 2)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    ct = pk.encrypt(pt, r)
    base = sk.decrypt(ct)
    monkeypatch.setenv("IPCLB200_DECRYPT", "tile")
    alt = sk.decrypt(ct)
    monkeypatch.delenv("IPCLB200_DECRYPT")
    assert np.array_equal(base, pt)
    assert np.array_equal(alt, pt)


@needs_experiments
@pytest.mark.parametrize("sq_layout,mode", [("1", "sqr"), ("2", "sqr2")])
def test_symmetric_squaring_kernel(capi, keys, monkeypatch, sq_layout, mode):
    """MontSqr::sqr / MontSqr2::sqr (mont_sqr.cuh, opt-in IPCLB200_DECRYPT=sqr|sqr2,
    16x4 and 32x2 lane layouts; k32 = the plain CIOS kernel at 32x2): single
    squarings are a^2 * 2^-2048 mod n below 2^2048 for edge values and random
    ones, and the decrypt kernel built on it leaves the same residues as the
    default kernel and as Python pow()"""
    rng = np.random.default_rng(4096)
    R = 1 << 2048
    mod = random_limbs(rng, 1, 64)
    mod[0, 0] |= 1
    mod[0, -1] |= 0x80000000
    n = from_limbs(mod[0])
    vals = [0, 1, R - 1, n - 1, n, int("ffffffff00000000" * 32, 16)] + \
        [int.from_bytes(rng.bytes(256), "little") for _ in range(70)]
    monkeypatch.setenv("IPCLB200_DEBUG_SQR_LAYOUT", sq_layout)
    s, m = capi.debug_montsqr(batch_to_limbs(vals, 64), mod[0])
    for got in (batch_from_limbs(s), batch_from_limbs(m)):
        assert all(x < R and (x * R - v * v) % n == 0 for x, v in zip(got, vals))
    k = keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    nsq = (p * q) ** 2
    count = 2500  # above the wide-layout threshold: the 16x4 kernels run
    cts = [int.from_bytes(rng.bytes(512), "little") % nsq for _ in range(count)]
    cts[:4] = [0, 1, nsq - 1, p * 999]
    ct = batch_to_limbs(cts, 128)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    monkeypatch.setenv("IPCLB200_DECRYPT", "int")
    base = sk.crt_residues(ct)
    monkeypatch.setenv("IPCLB200_DECRYPT", mode)
    x = sk.crt_residues(ct)
    assert np.array_equal(x, base)
    monkeypatch.setenv("IPCLB200_DECRYPT", "k32")
    assert np.array_equal(sk.crt_residues(ct), base)
    for i in range(40):
        assert from_limbs(x[i, 0]) == pow(cts[i], p - 1, p * p)
        assert from_limbs(x[i, 1]) == pow(cts[i], q - 1, q * q)


FP_CONFIGS = [
    ("fp", {"IPCLB200_FP_BLOCKS": "2"}),
    ("fp", {"IPCLB200_FP_BLOCKS": "3"}),
    ("dual", {"IPCLB200_FP_MASK": "4"}),
    ("dual", {"IPCLB200_FP_MASK": "6"}),
    ("dual2", {"IPCLB200_DUAL2": "2,1"}),
    ("dual2", {"IPCLB200_DUAL2": "1,2"}),
]


def test_decrypt_full_width_residues_vs_pow(capi, keys, monkeypatch):
    """the full-width kernel (IPCLB200_DECRYPT=int; the fallback of the two-digit
    decrypt for primes that do not fill their words) leaves the canonical residues
    ct^(p-1) mod p^2, ct^(q-1) mod q^2, incl. ciphertexts 0, 1, n^2-1 and multiples
    of p and q^2"""
    _residues_vs_pow(capi, keys, "int", {}, monkeypatch)


@needs_experiments
@pytest.mark.parametrize("mode,env", FP_CONFIGS,
                         ids=["-".join([m] + list(e.values())) for m, e in FP_CONFIGS])
def test_decrypt_pipes_residues_vs_pow(capi, keys, mode, env, monkeypatch):
    _residues_vs_pow(capi, keys, mode, env, monkeypatch)


def _residues_vs_pow(capi, keys, mode, env, monkeypatch):
    """every pipe configuration of the CRT-decrypt modexp (integer pipe, FP64
    pipe, both in one kernel, both as two kernels sharing the work counter)
    leaves the canonical residues ct^(p-1) mod p^2, ct^(q-1) mod q^2 -- checked
    against Python pow() -- and the same plaintexts; includes ciphertexts 0, 1,
    n^2-1 and multiples of p and q^2 (residue 0 / the modulus itself)"""
    k = keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    nsq = n * n
    rng = np.random.default_rng(2200)
    count = 203
    cts = [int.from_bytes(rng.bytes(512), "little") % nsq for _ in range(count)]
    cts[:5] = [0, 1, nsq - 1, p * 12345, q * q * 3 % nsq]
    ct = batch_to_limbs(cts, 128)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    monkeypatch.setenv("IPCLB200_DECRYPT", "int")
    base = sk.decrypt(ct)
    monkeypatch.setenv("IPCLB200_DECRYPT", mode)
    for a, b in env.items():
        monkeypatch.setenv(a, b)
    x = sk.crt_residues(ct)
    assert x.shape == (count, 2, 64)
    got = [[from_limbs(x[i, s]) for s in (0, 1)] for i in range(count)]
    want = [[pow(c, p - 1, p * p), pow(c, q - 1, q * q)] for c in cts]
    assert got == want
    assert np.array_equal(sk.decrypt(ct), base)
    # a real round trip on top
    pk = capi.PubKey(to_limbs(n, 64), to_limbs(k["hs"], 128), 1024)
    pt = random_limbs(rng, 77, 64, top_mask=0x3FFFFFFF)
    assert np.array_equal(sk.decrypt(pk.encrypt(pt, random_limbs(rng, 77, 32))), pt)


def test_concurrent_callers_through_cabi(capi, keys):
    """the C ABI is re-entrant: four host threads encrypt and decrypt on the
    same key objects at once (the reference calls encrypt/decrypt concurrently
    on one key, test/test_cryptography.cpp:45-57); ctypes drops the GIL"""
    import threading
    k = keys["1024"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = 32
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), 512)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(3):
                count = int(rng.integers(1, 300))
                pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
                r = random_limbs(rng, count, NL // 2)
                ct = pk.encrypt(pt, r)
                if not np.array_equal(sk.decrypt(ct), pt):
                    errors.append("round trip %d" % seed)
                m = capi.modmul(ct, ct, to_limbs(n * n, 2 * NL))
                d = batch_from_limbs(sk.decrypt(m))
                if d != [2 * x % n for x in batch_from_limbs(pt)]:
                    errors.append("ct+ct %d" % seed)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_full_size_batch_65536(capi, keys):
    """BASELINE.json configs[1]/[2] at full size (2048-bit key, 65536 elements),
    checked through size-independent properties: decrypt(encrypt(m)) == m for
    every element, decrypt(ct_a * ct_b mod n^2) == a + b, and the ciphertexts
    of two different randoms for the same plaintext differ"""
    k = keys["2048"]
    p, q = k["p"], k["q"]
    n = p * q
    NL, count = 64, 65536
    rng = np.random.default_rng(65536)
    a = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    b = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    ca = pk.encrypt(a, random_limbs(rng, count, 32))
    ca2 = pk.encrypt(a, random_limbs(rng, count, 32))
    cb = pk.encrypt(b, random_limbs(rng, count, 32))
    assert np.array_equal(sk.decrypt(ca), a)
    assert np.array_equal(sk.decrypt(ca2), a)
    assert not np.array_equal(ca, ca2)
    s = sk.decrypt(capi.modmul(ca, cb, to_limbs(n * n, 128)))
    # a, b < 2^2046 so a + b < n: plain multi-limb sum
    t = a.astype(np.uint64) + b.astype(np.uint64)
    want = np.zeros_like(a)
    carry = np.zeros(count, dtype=np.uint64)
    for j in range(NL):
        v = t[:, j] + carry
        want[:, j] = (v & 0xFFFFFFFF).astype(np.uint32)
        carry = v >> 32
    assert np.array_equal(s, want)
