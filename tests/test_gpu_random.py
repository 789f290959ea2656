"""GPU parity of the device-side generator of the DJN randoms
(chacha20_fill_kernel behind ipclb200_random_dev / batch_random / encrypt_drbg,
include/ipcl_b200.h) against oracle/chacha20_ref.py, which tests/
test_chacha_ref.py pins to RFC 8439; and of encrypt with device-drawn r against
the oracle's encrypt on the same r (ipcl/pub_key.cpp:51-110)."""
import os
import sys

import numpy as np
import pytest

from pailliercryptolib_b200.limbs import random_limbs, to_limbs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "oracle"))
import chacha20_ref as cc  # noqa: E402

pytestmark = pytest.mark.gpu


def _seed(rng):
    key = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32)
    nonce = rng.integers(0, 2**32, 3, dtype=np.uint64).astype(np.uint32)
    return key, nonce


@pytest.mark.parametrize("words,bits,count,first", [
    (32, 1024, 67, 0), (32, 1000, 5, 7), (16, 512, 33, 123456), (16, 481, 9, 1),
    (48, 1536, 11, 3), (5, 131, 40, 0), (1, 17, 300, 2), (64, 2048, 3, 2**20),
    (18, 576, 7, 5)])
def test_random_dev_matches_rfc8439_reference(capi, words, bits, count, first):
    import torch
    rng = np.random.default_rng(words * 1000 + bits)
    key, nonce = _seed(rng)
    d = torch.full((count, words), -1, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    capi.random_dev(d.data_ptr(), count, words, bits, key, nonce, first, s.cuda_stream)
    s.synchronize()
    got = d.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, cc.batch_randoms(key, nonce, count, words, bits, first))


def test_batch_random_and_bad_arguments(capi):
    rng = np.random.default_rng(3)
    key, nonce = _seed(rng)
    b = capi.Batch(1500, 32)
    b.random(1024, key, nonce)
    got = b.download()
    want = cc.batch_randoms(key, nonce, 64, 32, 1024, 0)
    assert np.array_equal(got[:64], want)
    tail = cc.batch_randoms(key, nonce, 4, 32, 1024, 1496)
    assert np.array_equal(got[1496:], tail)
    # all elements distinct, top bit of the batch not constant
    assert len({bytes(r) for r in got}) == 1500
    with pytest.raises(capi.IpclB200Error):
        b.random(1025, key, nonce)  # more bits than the stride holds
    with pytest.raises(capi.IpclB200Error):
        b.random(0, key, nonce)


@pytest.mark.parametrize("bits", ["1024", "2048"])
def test_encrypt_drbg_equals_encrypt_with_the_same_randoms(capi, oracle, keys, bits):
    k = keys[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    count = 301
    rng = np.random.default_rng(int(bits) + 1)
    key, nonce = _seed(rng)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    nl, hsl = to_limbs(n, NL), to_limbs(k["hs"], 2 * NL)
    pk = capi.PubKey(nl, hsl, int(bits) // 2)
    ct = pk.encrypt_drbg(pt, key, nonce)
    r = cc.batch_randoms(key, nonce, count, NL // 2, int(bits) // 2, 0)
    assert np.array_equal(ct, pk.encrypt(pt, r))
    assert np.array_equal(ct, oracle.encrypt(nl, hsl, pt, r))
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    assert np.array_equal(sk.decrypt(ct), pt)
    # a different key gives different ciphertexts of the same plaintexts
    key2 = key.copy()
    key2[0] ^= 1
    assert not np.array_equal(pk.encrypt_drbg(pt, key2, nonce), ct)
    # non-DJN keys have no device-drawn r
    with pytest.raises(capi.IpclB200Error):
        capi.PubKey(nl).encrypt_drbg(pt, key, nonce)
