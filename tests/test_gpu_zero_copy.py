"""Page-locked caller buffers used in place by the host-pointer entry points
(ipclb200_encrypt / ipclb200_decrypt, include/ipcl_b200.h: ipclb200_zero_copy_count):
the kernels read plaintexts / ciphertexts and write ciphertexts through the
mapped alias of the caller's memory.  Results must be the bits of the staged
path and of the oracle (ipcl/pub_key.cpp:51-110, ipcl/pri_key.cpp:114-157)."""
import numpy as np
import pytest

from pailliercryptolib_b200.limbs import random_limbs, to_limbs

pytestmark = pytest.mark.gpu


def _key(capi, keys, bits):
    k = keys[str(bits)]
    p, q = sorted((k["p"], k["q"]))
    nl = bits // 32
    n = p * q
    nlimbs, hs = to_limbs(n, nl), to_limbs(k["hs"], 2 * nl)
    pk = capi.PubKey(nlimbs, hs, bits // 2)
    sk = capi.PrivKey(to_limbs(p, nl // 2), to_limbs(q, nl // 2))
    return pk, sk, nlimbs, hs, to_limbs(p, nl // 2), to_limbs(q, nl // 2)


@pytest.mark.parametrize("bits,count", [(2048, 3001), (1024, 5000), (3072, 1500)])
def test_pinned_buffers_in_place_match_staged_and_oracle(capi, keys, oracle, monkeypatch,
                                                         bits, count):
    pk, sk, nlimbs, hs, pl, ql = _key(capi, keys, bits)
    nl = bits // 32
    rng = np.random.default_rng(bits + count)
    pt = random_limbs(rng, count, nl, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, nl // 2)
    monkeypatch.setenv("IPCLB200_ZERO_COPY", "7")   # all three operand classes
    monkeypatch.delenv("IPCLB200_DECRYPT", raising=False)
    # staged reference: pageable numpy arrays
    z0 = capi.zero_copy_count()
    ct_staged = pk.encrypt(pt, r)
    dt_staged = sk.decrypt(ct_staged)
    assert capi.zero_copy_count() == z0, "pageable buffers must be staged"
    assert np.array_equal(dt_staged, pt)
    S = 96
    assert np.array_equal(ct_staged[:S], oracle.encrypt(nlimbs, hs, pt[:S], r[:S]))
    # page-locked buffers: used in place (the views start inside the allocations)
    off = 7
    pt_pin = capi.pinned_empty((count + off, nl))
    ct_pin = capi.pinned_empty((count + off, 2 * nl))
    dt_pin = capi.pinned_empty((count + off, nl))
    pt_pin[off:] = pt
    ct_pin[:] = 0xA5A5A5A5
    dt_pin[:] = 0x5A5A5A5A
    pk.encrypt(pt_pin[off:], r, out=ct_pin[off:])
    assert capi.zero_copy_count() == z0 + 2, "plaintexts in, ciphertexts out: both in place"
    assert np.array_equal(ct_pin[off:], ct_staged)
    assert np.all(ct_pin[:off] == 0xA5A5A5A5), "wrote outside the batch"
    sk.decrypt(ct_pin[off:], out=dt_pin[off:])
    assert capi.zero_copy_count() == z0 + 3, "ciphertexts read in place"
    assert np.array_equal(dt_pin[off:], pt)
    assert np.all(dt_pin[:off] == 0x5A5A5A5A)
    assert np.array_equal(dt_pin[off:off + S], oracle.decrypt_crt(pl, ql, ct_pin[off:off + S]))
    # default: the encrypt operands in place, the decrypt input staged (faster)
    monkeypatch.delenv("IPCLB200_ZERO_COPY", raising=False)
    z1 = capi.zero_copy_count()
    ct_pin[:] = 0
    pk.encrypt(pt_pin[off:], r, out=ct_pin[off:])
    sk.decrypt(ct_pin[off:], out=dt_pin[off:])
    assert capi.zero_copy_count() == z1 + 2
    assert np.array_equal(ct_pin[off:], ct_staged) and np.array_equal(dt_pin[off:], pt)
    # switched off: same bits through the staging copies
    monkeypatch.setenv("IPCLB200_ZERO_COPY", "0")
    z1 = capi.zero_copy_count()
    ct_pin[:] = 0
    pk.encrypt(pt_pin[off:], r, out=ct_pin[off:])
    sk.decrypt(ct_pin[off:], out=dt_pin[off:])
    assert capi.zero_copy_count() == z1
    assert np.array_equal(ct_pin[off:], ct_staged) and np.array_equal(dt_pin[off:], pt)


def test_small_and_non_djn_batches_stay_staged(capi, keys, monkeypatch):
    """below 1024 elements per device, and for key types whose kernels touch the
    output more than once (non-DJN: (n*m+1) then the product with r^n), pinned
    buffers are staged like pageable ones -- and give the same results"""
    monkeypatch.setenv("IPCLB200_ZERO_COPY", "7")
    pk, sk, nlimbs, hs, pl, ql = _key(capi, keys, 2048)
    rng = np.random.default_rng(77)
    count = 300
    pt = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, 32)
    pt_pin = capi.pinned_empty((count, 64))
    ct_pin = capi.pinned_empty((count, 128))
    pt_pin[:] = pt
    z0 = capi.zero_copy_count()
    pk.encrypt(pt_pin, r, out=ct_pin)
    assert capi.zero_copy_count() == z0
    assert np.array_equal(ct_pin, pk.encrypt(pt, r))
    assert np.array_equal(sk.decrypt(ct_pin), pt)
    # non-DJN key, large batch
    pk2 = capi.PubKey(nlimbs)
    count = 2048
    pt = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    r2 = random_limbs(rng, count, 64, top_mask=0x3FFFFFFF)
    r2[:, 0] |= 1
    pt_pin = capi.pinned_empty((count, 64))
    ct_pin = capi.pinned_empty((count, 128))
    pt_pin[:] = pt
    z0 = capi.zero_copy_count()
    pk2.encrypt(pt_pin, r2, out=ct_pin)
    assert capi.zero_copy_count() == z0
    assert np.array_equal(ct_pin, pk2.encrypt(pt, r2))
    dt_pin = capi.pinned_empty((count, 64))
    sk.decrypt(ct_pin, out=dt_pin)     # CRT decrypt reads the pinned ciphertexts in place
    assert capi.zero_copy_count() == z0 + 1
    assert np.array_equal(dt_pin, pt)
