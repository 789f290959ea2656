"""The Python binding (pailliercryptolib_b200/ipcl_python.py, SURVEY.md section
8f row 4): same homomorphic identities the reference's operation tests check
(test/test_ops.cpp:126-608), through device-resident batches."""
import numpy as np
import pytest

from pailliercryptolib_b200 import capi
from pailliercryptolib_b200 import ipcl_python as ip


def test_no_device_raises_not_falls_back():
    """without an sm_100 device the binding fails loudly (no CPU path)"""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.IpclB200Error):
        ip.PaillierKeypair.generate_keypair(1024, True)


@pytest.mark.gpu
def test_iso_key_roundtrip_and_ops(iso, keys):
    k = keys["2048"]
    ip.context.initializeContext("GPU")
    pk = ip.PaillierPublicKey(k["p"] * k["q"], 2048, True, hs=k["hs"])
    sk = ip.PaillierPrivateKey(pk, k["p"], k["q"])
    n = pk.n
    rng = np.random.default_rng(5)
    a = [int.from_bytes(rng.bytes(200), "little") for _ in range(33)]
    b = [int.from_bytes(rng.bytes(200), "little") for _ in range(33)]
    kk = [int(x) for x in rng.integers(0, 1 << 32, size=33)]
    ca, cb = pk.encrypt(a), pk.encrypt(b)
    assert len(ca) == 33
    assert sk.decrypt(ca) == a
    assert sk.decrypt(ca + cb) == [(x + y) % n for x, y in zip(a, b)]
    assert sk.decrypt(ca + b) == [(x + y) % n for x, y in zip(a, b)]
    assert sk.decrypt(ca * kk) == [(x * y) % n for x, y in zip(a, kk)]
    assert sk.decrypt(ca * 3) == [(x * 3) % n for x in a]
    assert sk.decrypt(ca + cb[4]) == [(x + b[4]) % n for x in a]
    assert sk.decrypt((ca + cb) * kk + a) == [((x + y) * z + x) % n for x, y, z in zip(a, b, kk)]
    assert sk.decrypt(ca.sum()) == sum(a) % n
    assert sk.decrypt(ca[7]) == a[7]
    sk.enableCRT(False)
    assert sk.decrypt(cb) == b
    # the ISO known answer through this binding: c1 * c2 mod n^2 decrypts to m0 + m1
    c1 = ip.PaillierEncryptedNumber(pk, ip._DevBatch.from_numpy(
        capi._c(np.frombuffer(iso["c1"].to_bytes(512, "little"), dtype="<u4"))[None, :]))
    c2 = ip.PaillierEncryptedNumber(pk, ip._DevBatch.from_numpy(
        capi._c(np.frombuffer(iso["c2"].to_bytes(512, "little"), dtype="<u4"))[None, :]))
    assert (c1 + c2).ciphertexts() == [iso["c1c2"]]
    assert sk.decrypt(c1 + c2) == iso["m1m2"]
    with pytest.raises(ValueError):
        ca + pk.encrypt([1, 2])
    with pytest.raises(ValueError):
        ca * [-1] * 33


@pytest.mark.gpu
def test_generate_keypair_on_gpu():
    for djn in (True, False):
        pk, sk = ip.PaillierKeypair.generate_keypair(1024, djn)
        assert pk.n.bit_length() == 1024 and sk.p * sk.q == pk.n
        if djn:
            assert sk.p % 4 == 3 and sk.q % 4 == 3
        vals = [0, 1, pk.n - 1, 123456789]
        assert sk.decrypt(pk.encrypt(vals)) == vals
        assert sk.decrypt(pk.encrypt(41) + 1) == 42
    with pytest.raises(ValueError):
        ip.PaillierKeypair.generate_keypair(8192)
    with pytest.raises(ValueError):
        ip.PaillierKeypair.generate_keypair(1022)
