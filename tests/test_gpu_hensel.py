"""GPU parity of the two-digit (Hensel) CRT decrypt -- decrypt_hensel_kernel,
pailliercryptolib_b200/csrc/mont_hensel.cuh -- through the C ABI: against the
full-width kernel (IPCLB200_DECRYPT=int), against Python pow() restating
ipcl/pri_key.cpp:114-167, and against the oracle at the full batch size."""
import os
import sys

import numpy as np
import pytest

from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,
                                          random_limbs, to_limbs)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import dec_crt  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def all_keys(keys):
    from conftest import load_golden
    extra = {name: {k: int(v, 16) for k, v in d.items()}
             for name, d in load_golden("keys_extra.json").items()}
    return {**keys, **extra}


def _ciphertexts(rng, p, q, count):
    """valid ciphertexts (units mod n^2): random residues, the edge values 1,
    n+1, n^2-1, and unreduced inputs >= n^2 that fill all 4*pl words"""
    n = p * q
    nsq = n * n
    words = 4 * ((p.bit_length() + 31) // 32)
    top = 1 << (32 * words)
    cts = []
    while len(cts) < count:
        c = int.from_bytes(rng.bytes(words * 4), "little") % nsq
        if c % p and c % q:
            cts.append(c)
    cts[:3] = [1, n + 1, nsq - 1]
    k = (top - 1) // nsq
    cts[3] = cts[10] + (k - 1) * nsq if k > 1 else cts[10]
    cts[4] = top - 1 if (top - 1) % p and (top - 1) % q else cts[11]
    return cts, words


@pytest.mark.parametrize("name", ["1024", "2048", "3072", "4096", "2048_low", "2048_high"])
def test_hensel_decrypt_matches_full_width_and_pow(capi, all_keys, name, monkeypatch):
    k = all_keys[name]
    p, q = sorted((k["p"], k["q"]))
    pl = (p.bit_length() + 31) // 32
    rng = np.random.default_rng(sum(map(ord, name)))
    count = 700 if pl <= 32 else 300
    cts, words = _ciphertexts(rng, p, q, count)
    ct = batch_to_limbs(cts, words)
    sk = capi.PrivKey(to_limbs(p, pl), to_limbs(q, pl))
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    got = sk.decrypt(ct)
    monkeypatch.setenv("IPCLB200_DECRYPT", "int")
    base = sk.decrypt(ct)
    monkeypatch.delenv("IPCLB200_DECRYPT")
    assert np.array_equal(got, base)
    vals = batch_from_limbs(got)
    for i in list(range(12)) + [count - 1]:
        assert vals[i] == dec_crt(p, q, cts[i]), (name, i)
    assert vals[0] == 0 and vals[1] == 1


@pytest.mark.parametrize("count", [1, 5, 16, 17, 31, 100])
def test_hensel_ragged_small_batches(capi, all_keys, count, monkeypatch):
    k = all_keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    rng = np.random.default_rng(count)
    pk = capi.PubKey(to_limbs(n, 64), to_limbs(k["hs"], 128), 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    pt = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    ct = pk.encrypt(pt, random_limbs(rng, count, 32))
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    assert np.array_equal(sk.decrypt(ct), pt)


@pytest.mark.parametrize("rows", ["4", "16"])
def test_hensel_row_unroll_variants(capi, all_keys, rows, monkeypatch):
    k = all_keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    rng = np.random.default_rng(int(rows))
    cts, words = _ciphertexts(rng, p, q, 400)
    ct = batch_to_limbs(cts, words)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    base = sk.decrypt(ct)
    monkeypatch.setenv("IPCLB200_HENSEL_ROWS", rows)
    assert np.array_equal(sk.decrypt(ct), base)
    assert batch_from_limbs(base[:4]) == [dec_crt(p, q, c) for c in cts[:4]]


@pytest.mark.parametrize("name", ["2048", "2048_low", "2048_high"])
@pytest.mark.parametrize("layout", ["-2", "-1", "0", "1", "2"])
def test_hensel_layouts_agree_with_pow(capi, all_keys, name, layout, monkeypatch):
    """every lane layout of the two-digit decrypt at 32-word primes -- one task
    per thread (-2: compact staging, 12 warps per SM; -1: with the prefetch
    buffer), per 2, 4, 8 lanes -- on edge ciphertexts, a ragged count that leaves
    lanes of the last warp idle, constant schedule included"""
    k = all_keys[name]
    p, q = sorted((k["p"], k["q"]))
    rng = np.random.default_rng(sum(map(ord, name)) + 7)
    count = 333
    cts, words = _ciphertexts(rng, p, q, count)
    ct = batch_to_limbs(cts, words)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    monkeypatch.setenv("IPCLB200_HENSEL_SPREAD", layout)
    monkeypatch.setenv("IPCLB200_HENSEL_W64", "1")
    got = batch_from_limbs(sk.decrypt(ct))
    assert got == [dec_crt(p, q, c) for c in cts]
    sk.set_schedule(True)
    assert batch_from_limbs(sk.decrypt(ct)) == got


@pytest.mark.parametrize("count", [14337, 24577, 45057])
def test_ragged_large_batches_take_the_thread_per_task_layout(capi, all_keys, count):
    """batch sizes just past a boundary of the layout cost model that are not a
    multiple of the 32 tasks of a warp: the automatic choice is the
    thread-per-task kernel, the last warp of each side is partly idle; every
    plaintext must come back, a sample is checked against Python pow()"""
    assert capi.decrypt_layout(count, 32, 148) == -2
    k = all_keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    rng = np.random.default_rng(count)
    pk = capi.PubKey(to_limbs(n, 64), to_limbs(k["hs"], 128), 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    pt = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    ct = pk.encrypt(pt, random_limbs(rng, count, 32))
    got = sk.decrypt(ct)
    assert np.array_equal(got, pt)
    cts = batch_from_limbs(ct[-3:])
    assert batch_from_limbs(got[-3:]) == [dec_crt(p, q, c) for c in cts]


def test_full_batch_65536_bit_exact_vs_oracle(capi, oracle, keys):
    """BASELINE.json configs[1] at full size, every element compared bit for bit
    with the oracle (AVX512-IFMA mb8 restatement, itself pinned to the ISO KAT
    and to the scalar oracle in tests/test_oracle.py): 65536 ciphertexts from
    the wide comb table the bench times, 65536 plaintexts from the two-digit
    decrypt."""
    if not oracle.have_ifma():
        pytest.skip("host without AVX512-IFMA: the scalar oracle needs minutes for 65536")
    k = keys["2048"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL, count = 64, 65536
    rng = np.random.default_rng(20481)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, 32)
    nl, hsl = to_limbs(n, NL), to_limbs(k["hs"], 2 * NL)
    pk = capi.PubKey(nl, hsl, 1024)
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    ct = pk.encrypt(pt, r)
    ct2 = pk.encrypt(pt, r)  # second call: the upgraded (wide window) table
    want_ct = oracle.encrypt_mb8(nl, hsl, pt, r)
    assert np.array_equal(ct, want_ct)
    assert np.array_equal(ct2, want_ct)
    got = sk.decrypt(ct)
    assert np.array_equal(got, oracle.decrypt_crt_mb8(to_limbs(p, 32), to_limbs(q, 32), ct))
    assert np.array_equal(got, pt)


def test_3072_bit_key_4096_elements_vs_oracle(capi, oracle, keys):
    """BASELINE.json configs[3] key size, 4096 elements, every one compared with
    the oracle (scalar for the 6144-bit encrypt, which mbx_exp_mb8 cannot do)"""
    k = keys["3072"]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL, count = 96, 4096
    rng = np.random.default_rng(3072)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, 48)
    nl, hsl = to_limbs(n, NL), to_limbs(k["hs"], 2 * NL)
    pk = capi.PubKey(nl, hsl, 1536)
    sk = capi.PrivKey(to_limbs(p, 48), to_limbs(q, 48))
    ct = pk.encrypt(pt, r)
    assert np.array_equal(ct, oracle.encrypt(nl, hsl, pt, r))
    got = sk.decrypt(ct)
    assert np.array_equal(got, pt)
    assert np.array_equal(got, oracle.decrypt_crt(to_limbs(p, 48), to_limbs(q, 48), ct))


@pytest.mark.parametrize("name", ["1024", "2048", "3072", "4096", "2048_low"])
def test_constant_schedule_mode(capi, all_keys, name, monkeypatch):
    """ipclb200_privkey_set_schedule(1): fixed 4-bit windows over a table of all 16
    powers (the operation sequence no longer depends on the bits of p-1, q-1,
    lambda); same plaintexts as the sliding-window schedules, CRT and RAW"""
    k = all_keys[name]
    p, q = sorted((k["p"], k["q"]))
    pl = (p.bit_length() + 31) // 32
    rng = np.random.default_rng(len(name) + pl)
    cts, words = _ciphertexts(rng, p, q, 150)
    ct = batch_to_limbs(cts, words)
    sk = capi.PrivKey(to_limbs(p, pl), to_limbs(q, pl))
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    base = sk.decrypt(ct)
    base_raw = sk.decrypt(ct[:40], use_crt=False)
    sk.set_schedule(True)
    assert np.array_equal(sk.decrypt(ct), base)
    assert np.array_equal(sk.decrypt(ct[:40], use_crt=False), base_raw)
    sk.set_schedule(False)
    assert np.array_equal(sk.decrypt(ct), base)
    assert batch_from_limbs(base[:6]) == [dec_crt(p, q, c) for c in cts[:6]]


@pytest.mark.parametrize("name", ["1024", "2048", "3072", "4096", "2048_low", "2048_high"])
def test_hensel_encrypt_matches_oracle(capi, oracle, all_keys, name, monkeypatch):
    """encrypt_hensel_kernel (two-digit arithmetic mod n^2 over a fixed-base table
    of pairs) against the oracle, bit for bit: edge plaintexts 0, 1, n-1, edge
    randoms 0, 1, 2^k, all-ones, short randoms (fewer windows), narrow plaintext
    buffers; and the same ciphertexts from the full-width table.  2048_low has an
    n that does not fill its words and stays on the full-width kernel."""
    k = all_keys[name]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = (p.bit_length() + 31) // 32 * 2
    rb = NL * 16
    rng = np.random.default_rng(NL + len(name))
    count = 400
    pts = [int.from_bytes(rng.bytes(NL * 4), "little") % n for _ in range(count)]
    pts[:3] = [0, 1, n - 1]
    rs = [int.from_bytes(rng.bytes(NL * 2), "little") for _ in range(count)]
    rs[:5] = [0, 1, 1 << (rb - 1), (1 << rb) - 1, 1 << 16]
    pt, r = batch_to_limbs(pts, NL), batch_to_limbs(rs, NL // 2)
    nl, hsl = to_limbs(n, NL), to_limbs(k["hs"], 2 * NL)
    monkeypatch.setenv("IPCLB200_COMB_SYNC", "1")
    want = oracle.encrypt(nl, hsl, pt, r)

    def encrypt_with(policy_mb):
        pk = capi.PubKey(nl, hsl, rb)
        pk.set_table_policy(max_table_mb=policy_mb, upgrade_after=0)
        out = pk.encrypt(pt, r)
        short = pk.encrypt(pt[:64], np.ascontiguousarray(r[:64, :2]))   # 64-bit randoms
        narrow = pk.encrypt(np.ascontiguousarray(pt[:64, :1]), r[:64])  # 32-bit plaintexts
        pk.close()
        return out, short, narrow

    got, short, narrow = encrypt_with(48)
    assert np.array_equal(got, want)
    r_short = np.zeros((64, NL // 2), dtype=np.uint32)
    r_short[:, :2] = r[:64, :2]
    assert np.array_equal(short, oracle.encrypt(nl, hsl, pt[:64], r_short))
    pt_narrow = np.zeros((64, NL), dtype=np.uint32)
    pt_narrow[:, 0] = pt[:64, 0]
    assert np.array_equal(narrow, oracle.encrypt(nl, hsl, pt_narrow, r[:64]))
    monkeypatch.setenv("IPCLB200_NO_HENSEL_ENCRYPT", "1")
    got2, _, _ = encrypt_with(48)
    assert np.array_equal(got2, want)


@pytest.mark.parametrize("name", ["1024", "2048", "3072", "4096"])
def test_hensel_modexp_square_modulus_vs_oracle(capi, oracle, all_keys, name, monkeypatch):
    """modexp_hensel_kernel: a batch of a^b mod n^2 (CipherText * PlainText) when the
    modulus is the square of an n that fills its words.  Per-element exponents of
    1 ... 70 bits and of full width, a shared exponent, bases 0, 1, n^2-1 and
    unreduced ones up to 2^(32L)-1; against the oracle and against the full-width
    kernel (IPCLB200_NO_HENSEL_MODEXP=1)."""
    k = all_keys[name]
    n = k["p"] * k["q"]
    nsq = n * n
    L = (nsq.bit_length() + 31) // 32
    rng = np.random.default_rng(L)
    count = 2100
    top = 1 << (32 * L)
    bases = [int.from_bytes(rng.bytes(4 * L), "little") % nsq for _ in range(count)]
    bases[:6] = [0, 1, nsq - 1, n, top - 1, nsq + 12345]
    base = batch_to_limbs(bases, L)
    mod = to_limbs(nsq, L)
    # small exponents, every bit length from 1 to 70
    exps = [(int.from_bytes(rng.bytes(9), "little") >> (72 - 1 - i % 70)) | 1 for i in range(count)]
    exps[:3] = [0, 1, 2]
    e_small = batch_to_limbs(exps, 3)
    monkeypatch.delenv("IPCLB200_NO_HENSEL_MODEXP", raising=False)
    got = capi.modexp(base, e_small, mod, capi.SHARED_MOD)
    want = oracle.modexp(base, e_small, mod[None, :], shared_mod=True)
    assert np.array_equal(got, want)
    assert batch_from_limbs(got[:8]) == [pow(b, e, nsq) for b, e in zip(bases[:8], exps[:8])]
    # shared exponent
    e_sh = to_limbs(0xC0FFEE1234567, 2)
    got_sh = capi.modexp(base, e_sh, mod, capi.SHARED_MOD | capi.SHARED_EXP)
    assert batch_from_limbs(got_sh[:40]) == [pow(b, 0xC0FFEE1234567, nsq) for b in bases[:40]]
    monkeypatch.setenv("IPCLB200_NO_HENSEL_MODEXP", "1")
    assert np.array_equal(capi.modexp(base, e_small, mod, capi.SHARED_MOD), want)
    assert np.array_equal(capi.modexp(base, e_sh, mod, capi.SHARED_MOD | capi.SHARED_EXP), got_sh)
    monkeypatch.delenv("IPCLB200_NO_HENSEL_MODEXP")
    # full-width exponents (the benchmark shape of ct*pt), a sample against pow()
    e_full = random_limbs(rng, count, L // 2)
    got_full = capi.modexp(base, e_full, mod, capi.SHARED_MOD)
    ef = batch_from_limbs(e_full[:10])
    assert batch_from_limbs(got_full[:10]) == [pow(b, e, nsq) for b, e in zip(bases[:10], ef)]
    monkeypatch.setenv("IPCLB200_NO_HENSEL_MODEXP", "1")
    assert np.array_equal(capi.modexp(base, e_full, mod, capi.SHARED_MOD), got_full)
