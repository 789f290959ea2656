"""The multi-GPU surface of the C ABI (include/ipcl_b200.h): batches sharded over
the devices of ipclb200_init_devices, host-pointer calls split over them, NCCL
scatter/gather, and the lock-free *_dev entry points on concurrent streams.
Everything is compared bit for bit with the oracle.  Cases that need more than
one GPU skip on a one-GPU box (the sharded code path still runs with one shard)."""
import threading

import numpy as np
import pytest

from pailliercryptolib_b200.limbs import random_limbs, to_limbs

pytestmark = pytest.mark.gpu


def _key(capi, keys, bits="2048"):
    k = keys[bits]
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    nl, hsl = to_limbs(n, NL), to_limbs(k["hs"], 2 * NL)
    pk = capi.PubKey(nl, hsl, int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    return p, q, n, NL, nl, hsl, pk, sk


@pytest.fixture
def all_devices(capi):
    n = capi.init_devices(0)
    yield n
    capi.init_devices(1)


def test_batch_pipeline_matches_oracle(capi, oracle, keys, all_devices):
    """encrypt -> ct+ct -> ct*k -> decrypt on sharded batches; on an N-GPU box the
    5000 elements are split into N contiguous blocks"""
    p, q, n, NL, nl, hsl, pk, sk = _key(capi, keys)
    rng = np.random.default_rng(77)
    count = 5000
    a = random_limbs(rng, count, NL, top_mask=0x0FFFFFFF)
    b = random_limbs(rng, count, NL, top_mask=0x0FFFFFFF)
    ra, rb = random_limbs(rng, count, 32), random_limbs(rng, count, 32)
    nsq = to_limbs(n * n, 2 * NL)
    A, B = capi.Batch.from_host(a), capi.Batch.from_host(b)
    RA, RB = capi.Batch.from_host(ra), capi.Batch.from_host(rb)
    assert len(A.shards()) == min(all_devices, count // 512)
    CA, CB, S = capi.Batch(count, 2 * NL), capi.Batch(count, 2 * NL), capi.Batch(count, 2 * NL)
    pk.encrypt_batch(A, RA, CA)
    pk.encrypt_batch(B, RB, CB)
    capi.modmul_batch(CA, CB, nsq, S)
    ca = CA.download()
    assert np.array_equal(ca, oracle.encrypt(nl, hsl, a, ra))
    s = S.download()
    assert np.array_equal(s, oracle.modmul(ca, CB.download(), nsq))
    # ct * 3 (shared exponent) and ct + one shared ciphertext
    M = capi.Batch(count, 2 * NL)
    capi.modexp_batch(S, None, nsq, M, exp_shared=np.array([3], dtype=np.uint32))
    T = capi.Batch(count, 2 * NL)
    capi.modmul_batch(M, None, nsq, T, b_shared=ca[0])
    PT = capi.Batch(count, NL)
    sk.decrypt_batch(T, PT)
    got = PT.download()
    want = oracle.decrypt_crt(to_limbs(p, 32), to_limbs(q, 32), T.download())
    assert np.array_equal(got, want)
    from pailliercryptolib_b200.limbs import batch_from_limbs
    av, bv, gv = (batch_from_limbs(x[:50]) for x in (a, b, got))
    assert gv == [(3 * (x + y) + av[0]) % n for x, y in zip(av, bv)]


def test_host_pointer_calls_split_over_devices(capi, oracle, keys, all_devices):
    p, q, n, NL, nl, hsl, pk, sk = _key(capi, keys)
    rng = np.random.default_rng(78)
    count = 4099  # not a multiple of the device count
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, 32)
    ct = pk.encrypt(pt, r)
    assert np.array_equal(ct, oracle.encrypt(nl, hsl, pt, r))
    assert np.array_equal(sk.decrypt(ct), pt)
    assert np.array_equal(sk.decrypt(ct, use_crt=False), pt)
    nsq = to_limbs(n * n, 2 * NL)
    assert np.array_equal(capi.modmul(ct, ct[::-1].copy(), nsq),
                          oracle.modmul(ct, ct[::-1].copy(), nsq))
    e = random_limbs(rng, count, 2)
    got = capi.modexp(ct, e, nsq, capi.SHARED_MOD)
    assert np.array_equal(got, oracle.modexp(ct, e, nsq[None, :], shared_mod=True))


def test_scatter_gather_roundtrip(capi, keys, all_devices):
    """ipclb200_batch_scatter / gather: one contiguous buffer on the first device
    <-> shards (NCCL send/recv when there is more than one shard)"""
    import torch
    count, words = 6000, 128
    rng = np.random.default_rng(5)
    src = random_limbs(rng, count, words)
    d_src = torch.from_numpy(src.view(np.int32)).to("cuda:0")
    d_dst = torch.zeros_like(d_src)
    torch.cuda.synchronize()
    b = capi.Batch(count, words)
    b.scatter_from(d_src.data_ptr())
    assert np.array_equal(b.download(), src)
    b.gather_to(d_dst.data_ptr())
    b.sync()
    capi.lib().ipclb200_sync()
    torch.cuda.synchronize()
    assert np.array_equal(d_dst.cpu().numpy().view(np.uint32), src)


def test_two_devices_from_one_process(capi, oracle, keys):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    assert capi.init_devices(2) == 2
    try:
        p, q, n, NL, nl, hsl, pk, sk = _key(capi, keys)
        rng = np.random.default_rng(79)
        count = 3000
        pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
        r = random_limbs(rng, count, 32)
        A, R, C = capi.Batch.from_host(pt), capi.Batch.from_host(r), capi.Batch(count, 2 * NL)
        sh = C.shards()
        assert [s["device"] for s in sh] == [0, 1] and sh[0]["count"] + sh[1]["count"] == count
        pk.encrypt_batch(A, R, C)
        assert np.array_equal(C.download(), oracle.encrypt(nl, hsl, pt, r))
        P = capi.Batch(count, NL)
        sk.decrypt_batch(C, P)
        assert np.array_equal(P.download(), pt)
    finally:
        capi.init_devices(1)


def test_dev_calls_on_concurrent_streams(capi, keys):
    """two decrypt_dev / encrypt_dev chains enqueued on two caller streams at the
    same time share no scratch: both give the right plaintexts (round 1 kept the
    CRT residues in one global buffer)"""
    import torch
    p, q, n, NL, nl, hsl, pk, sk = _key(capi, keys)
    rng = np.random.default_rng(80)
    count = 3000
    pts = [random_limbs(rng, count, NL, top_mask=0x3FFFFFFF) for _ in range(2)]
    rs = [random_limbs(rng, count, 32) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    d_pt = [torch.from_numpy(x.view(np.int32)).cuda() for x in pts]
    d_r = [torch.from_numpy(x.view(np.int32)).cuda() for x in rs]
    d_ct = [torch.zeros((count, 2 * NL), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_out = [torch.zeros((count, NL), dtype=torch.int32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(2):
            s = streams[i].cuda_stream
            pk.encrypt_dev(d_pt[i].data_ptr(), NL, d_r[i].data_ptr(), 32, count,
                           d_ct[i].data_ptr(), s)
            sk.decrypt_dev(d_ct[i].data_ptr(), count, d_out[i].data_ptr(), s)
    for s in streams:
        s.synchronize()
    for i in range(2):
        assert np.array_equal(d_out[i].cpu().numpy().view(np.uint32), pts[i])


def test_four_host_threads_overlap(capi, keys):
    """the reference's contract: 4 threads on one key (test_cryptography.cpp:45-57).
    Every call runs on its own stream; results stay exact."""
    p, q, n, NL, nl, hsl, pk, sk = _key(capi, keys, "1024")
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(4):
                count = 256
                pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
                ct = pk.encrypt(pt, random_limbs(rng, count, NL // 2))
                if not np.array_equal(sk.decrypt(ct), pt):
                    errors.append(seed)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
