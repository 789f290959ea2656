#!/usr/bin/env python3
"""The other BASELINE.json configs (bench.py is configs[1], the headline):

  configs[2]  2048-bit key, batch 65536: HE add (ct+ct) and mul (ct*pt)
  configs[3]  3072-bit key, batch 262144 encrypt+decrypt, sharded over the
              ranks with an NCCL scatter before and gather after
  configs[4]  raw modexp microbench, 1024/2048/3072/4096-bit, batch 2^10..2^20

One JSON line per measurement (rank 0).  Device-resident inputs, CUDA events on
the launching stream, >= 3 warm-ups, max over ranks.  Every timed result is
also checked: homomorphic identity / round trip / oracle sample.

  python benchmarks/bench_configs.py [--only add,mul,raw,key3072]
  torchrun --nproc-per-node N benchmarks/bench_configs.py --only key3072
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pailliercryptolib_b200 import capi, sharding  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402


def mont_macs(L):
    return 2 * L * L + L


def modexp_macs(L, e_bits):
    return (e_bits + (e_bits + 4) // 5 + 34) * mont_macs(L)


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms))


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(device)


def host(t):
    return t.cpu().numpy().view(np.uint32)


def load_keys():
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        return {b: {k: int(v, 16) for k, v in d.items()} for b, d in json.load(f).items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="add,mul,raw,key3072,cpu")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--raw-max-log2", type=int, default=20)
    ap.add_argument("--batch3072", type=int, default=262144)
    args = ap.parse_args()
    only = set(args.only.split(","))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    capi.init(local)
    stream = torch.cuda.current_stream().cuda_stream
    peak, _ = capi.int_peak()
    keys = load_keys()
    import oracle as orc

    def emit(d):
        if rank == 0:
            d["n_gpus"] = world
            d["int_peak_tmac32"] = peak / 1e12
            print(json.dumps(d), flush=True)

    rng = np.random.default_rng(0xB200 + rank)

    if ("add" in only or "mul" in only) and world == 1:
        k = keys["2048"]
        p, q = sorted((k["p"], k["q"]))
        n = p * q
        NL, B = 64, 65536
        nsq = to_limbs(n * n, 2 * NL)
        pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), 1024)
        sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
        a = random_limbs(rng, B, NL, top_mask=0x3FFFFFFF)
        b = random_limbs(rng, B, NL, top_mask=0x3FFFFFFF)
        d_a, d_b = dev(a, device), dev(b, device)
        d_r1, d_r2 = dev(random_limbs(rng, B, 32), device), dev(random_limbs(rng, B, 32), device)
        d_ca = torch.empty((B, 2 * NL), dtype=torch.int32, device=device)
        d_cb = torch.empty_like(d_ca)
        d_out = torch.empty_like(d_ca)
        d_dt = torch.empty((B, NL), dtype=torch.int32, device=device)
        pk.encrypt_dev(d_a.data_ptr(), NL, d_r1.data_ptr(), 32, B, d_ca.data_ptr(), stream)
        pk.encrypt_dev(d_b.data_ptr(), NL, d_r2.data_ptr(), 32, B, d_cb.data_ptr(), stream)
        if "add" in only:
            ms = timed(lambda: capi.modmul_dev(d_ca.data_ptr(), d_cb.data_ptr(), nsq, B,
                                               d_out.data_ptr(), stream), args.steps)
            sk.decrypt_dev(d_out.data_ptr(), B, d_dt.data_ptr(), stream)
            torch.cuda.synchronize()
            got = host(d_dt)
            # a, b < 2^2046 so a+b < n: plain integer sum
            s = (a.astype(np.uint64) + b.astype(np.uint64))
            carry = np.zeros(B, dtype=np.uint64)
            want = np.zeros_like(a)
            for j in range(NL):
                t = s[:, j] + carry
                want[:, j] = (t & 0xFFFFFFFF).astype(np.uint32)
                carry = t >> 32
            assert np.array_equal(got, want), "dec(ct+ct) != a+b"
            emit({"config": "2048-bit key, batch=65536 HE add (ct+ct)", "ops_per_s": B / ms * 1e3,
                  "ms": ms, "kernel": "modmul_kernel<16,8>",
                  "alg_mac32_per_op": 2 * mont_macs(128),
                  "roofline_frac_int": B * 2 * mont_macs(128) / (ms * 1e-3) / peak,
                  "alg_bytes_per_op": 1536, "hbm_gbs": B * 1536 / (ms * 1e-3) / 1e9,
                  "verified": "decrypt(sum) == a + b for all 65536"})
        if "mul" in only:
            for ebits, words in ((2048, 64), (32, 1)):
                e = random_limbs(rng, B, words)
                d_e = dev(e, device)
                ms = timed(lambda: capi.modexp_dev(d_ca.data_ptr(), d_e.data_ptr(), nsq, words,
                                                   ebits, B, d_out.data_ptr(), stream),
                           args.steps)
                torch.cuda.synchronize()
                sample = slice(0, 16)
                want = orc.modexp(host(d_ca)[sample], e[sample], nsq[None, :], shared_mod=True)
                assert np.array_equal(host(d_out)[sample], want), "ct*pt mismatch vs oracle"
                emit({"config": "2048-bit key, batch=65536 HE mul (ct*pt), %d-bit plaintext" % ebits,
                      "ops_per_s": B / ms * 1e3, "ms": ms, "kernel": "modexp_kernel<16,8>",
                      "alg_mac32_per_op": modexp_macs(128, ebits),
                      "roofline_frac_int": B * modexp_macs(128, ebits) / (ms * 1e-3) / peak,
                      "verified": "first 16 vs oracle"})

    if "raw" in only and world == 1:
        for bits in (1024, 2048, 3072, 4096):
            L = bits // 32
            mod = random_limbs(rng, 1, L)
            mod[0, 0] |= 1
            mod[0, -1] |= 0x80000000
            for lg in range(10, args.raw_max_log2 + 1, 2):
                B = 1 << lg
                if bits >= 3072 and lg > 18:
                    continue
                base, exp = random_limbs(rng, B, L), random_limbs(rng, B, L)
                d_b, d_e = dev(base, device), dev(exp, device)
                d_o = torch.empty_like(d_b)
                ms = timed(lambda: capi.modexp_dev(d_b.data_ptr(), d_e.data_ptr(), mod[0], L,
                                                   bits, B, d_o.data_ptr(), stream),
                           max(1, args.steps - 1), warmup=3 if lg <= 14 else 1)
                torch.cuda.synchronize()
                want = orc.modexp(base[:4], exp[:4], mod, shared_mod=True)
                assert np.array_equal(host(d_o)[:4], want)
                emit({"config": "raw modexp %d-bit modulus, %d-bit exponent, batch 2^%d" % (bits, bits, lg),
                      "modexp_per_s": B / ms * 1e3, "ms": ms,
                      "alg_mac32_per_op": modexp_macs(L, bits),
                      "roofline_frac_int": B * modexp_macs(L, bits) / (ms * 1e-3) / peak,
                      "hbm_gbs": B * 3 * L * 4 / (ms * 1e-3) / 1e9, "verified": "first 4 vs oracle"})

    if "rawmg" in only:
        # configs[4] across the GPUs of one box: a fixed total batch is cut into
        # contiguous shards (strong scaling, no collective on the data path:
        # every modexp is independent); time = max over ranks
        for bits, lgs in ((1024, (10, 14, 20)), (2048, (10, 14, 18, 20)), (3072, (14, 18)),
                          (4096, (16,))):
            L = bits // 32
            g = np.random.default_rng(bits)
            mod = random_limbs(g, 1, L)
            mod[0, 0] |= 1
            mod[0, -1] |= 0x80000000
            for lg in lgs:
                total = 1 << lg
                s0, s1 = sharding.shard_range(total, world, rank)
                Bl = s1 - s0
                gl = np.random.default_rng(bits * 100 + lg * 10 + rank)
                base, exp = random_limbs(gl, Bl, L), random_limbs(gl, Bl, L)
                d_b, d_e = dev(base, device), dev(exp, device)
                d_o = torch.empty_like(d_b)
                fn = lambda: capi.modexp_dev(d_b.data_ptr(), d_e.data_ptr(), mod[0], L, bits, Bl,
                                             d_o.data_ptr(), stream)
                fn()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                ms = timed(fn, 2, warmup=1)
                ms = sharding.max_over_ranks([ms], device)[0]
                want = orc.modexp(base[:2], exp[:2], mod, shared_mod=True)
                assert np.array_equal(host(d_o)[:2], want)
                emit({"config": "raw modexp %d-bit, total batch 2^%d sharded over the ranks" % (bits, lg),
                      "modexp_per_s": total / ms * 1e3, "ms": ms, "batch_per_gpu": Bl,
                      "alg_mac32_per_op": modexp_macs(L, bits),
                      "roofline_frac_int_per_gpu": total * modexp_macs(L, bits) / (ms * 1e-3) / peak / world,
                      "hbm_gbs": total * 3 * L * 4 / (ms * 1e-3) / 1e9,
                      "verified": "first 2 of every rank vs oracle"})

    if "cpu" in only and rank == 0:
        # CPU lines on this box's host cores, same inputs for the three
        # implementations: scalar port, restated AVX512-IFMA mb8, OpenSSL
        import time
        for bits, ebits, count in ((2048, 1024, 2048), (4096, 1024, 1024)):
            L, EL = bits // 32, ebits // 32
            mod = random_limbs(rng, 1, L)
            mod[0, 0] |= 1
            mod[0, -1] |= 0x80000000
            base, exp = random_limbs(rng, count, L), random_limbs(rng, count, EL)
            impls = [("scalar radix-2^32 port", lambda: orc.modexp(base, exp, mod, shared_mod=True))]
            if orc.have_ifma():
                impls.append(("AVX512-IFMA mb8 (restated mbx_exp_mb8)",
                              lambda: orc.modexp_mb8(base, exp, mod[0])))
            if orc.have_openssl():
                impls.append(("OpenSSL BN_mod_exp_mont_consttime",
                              lambda: orc.modexp_openssl(base, exp, mod[0])))
            ref = None
            for name, fn in impls:
                fn()
                t0 = time.perf_counter()
                out = fn()
                dt = time.perf_counter() - t0
                ref = out if ref is None else ref
                assert np.array_equal(out, ref)
                emit({"config": "CPU modexp %d-bit modulus, %d-bit exponent, %d elements" % (bits, ebits, count),
                      "impl": name, "threads": orc.num_threads(), "modexp_per_s": count / dt})

    if "key3072" in only:
        k = keys["3072"]
        p, q = sorted((k["p"], k["q"]))
        n = p * q
        NL = 96
        total = args.batch3072
        pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), 1536)
        sk = capi.PrivKey(to_limbs(p, 48), to_limbs(q, 48))
        full_pt = full_r = None
        if rank == 0:
            g = np.random.default_rng(3072)
            full_pt = dev(random_limbs(g, total, NL, top_mask=0x3FFFFFFF), device)
            full_r = dev(random_limbs(g, total, 48), device)
        s0, s1 = sharding.shard_range(total, world, rank)
        Bl = s1 - s0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        result = {}

        def run():
            ev[0].record()
            pt = sharding.scatter_rows(full_pt, total, NL, torch.int32, device)
            r = sharding.scatter_rows(full_r, total, 48, torch.int32, device)
            ev[1].record()
            ct = torch.empty((Bl, 2 * NL), dtype=torch.int32, device=device)
            dt = torch.empty((Bl, NL), dtype=torch.int32, device=device)
            pk.encrypt_dev(pt.data_ptr(), NL, r.data_ptr(), 48, Bl, ct.data_ptr(), stream)
            sk.decrypt_dev(ct.data_ptr(), Bl, dt.data_ptr(), stream)
            ev[2].record()
            out = sharding.gather_rows(dt, total)
            ev[3].record()
            ev[3].synchronize()
            result["out"] = out

        for _ in range(2):
            run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        run()
        t = sharding.max_over_ranks([ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]),
                                     ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], device)
        if rank == 0:
            assert torch.equal(result["out"], full_pt), "3072-bit round trip failed"
        macs = modexp_macs(192, 1536) + 2 * modexp_macs(96, 1536)
        emit({"config": "3072-bit key, batch=%d encrypt+decrypt (DJN, CRT), scatter -> shard -> gather" % total,
              "pairs_per_s": total / t[0] * 1e3, "ms_total": t[0], "ms_scatter": t[1],
              "ms_compute": t[2], "ms_gather": t[3], "alg_mac32_per_pair": macs,
              "roofline_frac_int_per_gpu": total * macs / (t[2] * 1e-3) / peak / world,
              "verified": "decrypt(encrypt(pt)) == pt for the whole batch on rank 0"})

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
