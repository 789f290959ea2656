"""Pins the CPU oracle (oracle/paillier_oracle.c) before anything trusts it:
against the reference's own ISO/IEC 18033-6 known-answer vectors
(/root/reference/test/test_cryptography.cpp:104-241) and against the committed
Python-pow golden vectors."""
import math

import numpy as np

from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,
                                          from_limbs, to_limbs)


def test_iso_kat_encrypt_add_decrypt(oracle, iso):
    """enc(m0;r0)=c1, enc(m1;r1)=c2, c1*c2 mod n^2 = c1c2, dec(c1c2)=m0+m1 --
    the assertions of ISO_IEC_18033_6_ComplianceTest, test_cryptography.cpp:217-240."""
    n = iso["p"] * iso["q"]
    NL = 64
    pt = batch_to_limbs([iso["m0"], iso["m1"]], NL)
    r = batch_to_limbs([iso["r0"], iso["r1"]], NL)
    ct = oracle.encrypt(to_limbs(n, NL), None, pt, r)
    assert from_limbs(ct[0]) == iso["c1"]
    assert from_limbs(ct[1]) == iso["c2"]
    s = oracle.modmul(ct[0:1], ct[1:2], to_limbs(n * n, 2 * NL))
    assert from_limbs(s[0]) == iso["c1c2"]
    p, q = sorted((iso["p"], iso["q"]))
    d = oracle.decrypt_crt(to_limbs(p, 32), to_limbs(q, 32), np.vstack([ct, s]))
    assert batch_from_limbs(d) == [iso["m0"], iso["m1"], iso["m1m2"]]


def test_iso_kat_batch_of_21(oracle, iso):
    """the reference encrypts 21 values (chunks 8+8+5), test_cryptography.cpp:102,198-212"""
    n = iso["p"] * iso["q"]
    ms = [iso["m0"]] * 21
    rs = [iso["r0"]] * 21
    ms[1], rs[1] = iso["m1"], iso["r1"]
    ct = oracle.encrypt(to_limbs(n, 64), None, batch_to_limbs(ms, 64),
                        batch_to_limbs(rs, 64))
    got = batch_from_limbs(ct)
    assert got[0] == iso["c1"] and got[1] == iso["c2"]
    assert all(g == iso["c1"] for g in got[2:])
    p, q = sorted((iso["p"], iso["q"]))
    d = oracle.decrypt_crt(to_limbs(p, 32), to_limbs(q, 32), ct)
    assert batch_from_limbs(d) == ms


def test_decrypt_raw_matches(oracle, iso):
    p, q = iso["p"], iso["q"]
    n = p * q
    lam = math.lcm(p - 1, q - 1)
    x = pow((pow(n + 1, lam, n * n) - 1) // n, -1, n)
    ct = batch_to_limbs([iso["c1"], iso["c2"], iso["c1c2"]], 128)
    d = oracle.decrypt_raw(to_limbs(n, 64), to_limbs(lam, 64), to_limbs(x, 64), ct)
    assert batch_from_limbs(d) == [iso["m0"], iso["m1"], iso["m1m2"]]


def test_modexp_golden(oracle, modexp_vectors):
    by_bits = {}
    for v in modexp_vectors:
        by_bits.setdefault(v["bits"], []).append(v)
    for bits, vs in by_bits.items():
        L = bits // 32
        b = batch_to_limbs([int(v["b"], 16) for v in vs], L)
        e = batch_to_limbs([int(v["e"], 16) for v in vs], L)
        m = batch_to_limbs([int(v["m"], 16) for v in vs], L)
        r = oracle.modexp(b, e, m)
        assert batch_from_limbs(r) == [int(v["r"], 16) for v in vs], bits


def test_modexp_rejects_even_and_zero_modulus(oracle):
    import pytest
    b = batch_to_limbs([3], 16)
    e = batch_to_limbs([5], 16)
    with pytest.raises(ValueError):
        oracle.modexp(b, e, batch_to_limbs([10], 16))
    with pytest.raises(ValueError):
        oracle.modexp(b, e, batch_to_limbs([0], 16))


def test_scheme_golden(oracle, keys, scheme_vectors):
    for bits, items in scheme_vectors.items():
        k = keys[bits]
        p, q = sorted((k["p"], k["q"]))
        n = p * q
        NL = int(bits) // 32
        ms = [int(i["m"], 16) for i in items]
        pt = batch_to_limbs(ms, NL)
        r_djn = batch_to_limbs([int(i["r_djn"], 16) for i in items], NL // 2)
        r_std = batch_to_limbs([int(i["r_std"], 16) for i in items], NL)
        nl = to_limbs(n, NL)
        c_djn = oracle.encrypt(nl, to_limbs(k["hs"], 2 * NL), pt, r_djn)
        c_std = oracle.encrypt(nl, None, pt, r_std)
        c_plain = oracle.encrypt(nl, None, pt, None, make_secure=False)
        assert batch_from_limbs(c_djn) == [int(i["c_djn"], 16) for i in items]
        assert batch_from_limbs(c_std) == [int(i["c_std"], 16) for i in items]
        assert batch_from_limbs(c_plain) == [int(i["c_plain"], 16) for i in items]
        d = oracle.decrypt_crt(to_limbs(p, NL // 2), to_limbs(q, NL // 2), c_djn)
        assert batch_from_limbs(d) == ms
        nsq = to_limbs(n * n, 2 * NL)
        add = oracle.modmul(c_djn, c_std, nsq)
        assert batch_from_limbs(add) == [int(i["c_add"], 16) for i in items]
        kk = batch_to_limbs([int(i["k"], 16) for i in items], 2 * NL)
        mul = oracle.modexp(c_djn, kk, nsq[None, :], shared_mod=True)
        assert batch_from_limbs(mul) == [int(i["c_mul"], 16) for i in items]


def test_mb8_matches_scalar(oracle, iso, keys):
    """the AVX512-IFMA multi-buffer restatement (CPU baseline) is bit-exact
    with the scalar oracle and reproduces the ISO known answers"""
    import pytest
    if not oracle.have_ifma():
        pytest.skip("host has no AVX512-IFMA")
    rng = np.random.default_rng(5)
    from pailliercryptolib_b200.limbs import random_limbs
    for L, EL, count in [(32, 16, 19), (64, 32, 21), (96, 8, 9), (128, 32, 13)]:
        mod = random_limbs(rng, 1, L)
        mod[0, 0] |= 1
        mod[0, -1] |= 0x80000000
        base = random_limbs(rng, count, L)
        base[0] = 0
        base[1] = 0
        base[1, 0] = 1
        exp = random_limbs(rng, count, EL)
        exp[2] = 0
        want = oracle.modexp(base, exp, mod, shared_mod=True)
        got = oracle.modexp_mb8(base, exp, mod[0])
        assert np.array_equal(got, want), L
    n = iso["p"] * iso["q"]
    ms = [iso["m0"], iso["m1"]] + [iso["m0"]] * 7
    rs = [iso["r0"], iso["r1"]] + [iso["r0"]] * 7
    ct = oracle.encrypt_mb8(to_limbs(n, 64), None, batch_to_limbs(ms, 64),
                            batch_to_limbs(rs, 64))
    assert from_limbs(ct[0]) == iso["c1"] and from_limbs(ct[1]) == iso["c2"]
    p, q = sorted((iso["p"], iso["q"]))
    d = oracle.decrypt_crt_mb8(to_limbs(p, 32), to_limbs(q, 32), ct)
    assert batch_from_limbs(d) == ms
    k = keys["2048"]
    hs = to_limbs(k["hs"], 128)
    pt = random_limbs(rng, 11, 64, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, 11, 32)
    assert np.array_equal(oracle.encrypt_mb8(to_limbs(n, 64), hs, pt, r),
                          oracle.encrypt(to_limbs(n, 64), hs, pt, r))


def test_oracle_agrees_with_openssl(oracle):
    """third opinion: OpenSSL BN_mod_exp_mont_consttime (what the reference's
    own QAT tests compare against, module/heqat/test/test_bnModExp.cpp:60,205)"""
    import pytest
    if not oracle.have_openssl():
        pytest.skip("OpenSSL headers not installed")
    from pailliercryptolib_b200.limbs import random_limbs
    rng = np.random.default_rng(8)
    for L, EL, count in [(32, 32, 24), (64, 32, 16), (128, 64, 6), (192, 48, 3)]:
        mod = random_limbs(rng, 1, L)
        mod[0, 0] |= 1
        base = random_limbs(rng, count, L)
        exp = random_limbs(rng, count, EL)
        want = oracle.modexp_openssl(base, exp, mod[0])
        assert np.array_equal(oracle.modexp(base, exp, mod, shared_mod=True), want)
