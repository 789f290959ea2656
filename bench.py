#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json:
Paillier encrypt+decrypt throughput at a 2048-bit key on a 65536-element batch.

One "step" = encrypt 65536 plaintexts (DJN obfuscator hs^r, r = 1024 bit, the
default key type of ipcl::generateKeypair) and CRT-decrypt the 65536
ciphertexts that came out.  Key = the ISO/IEC 18033-6 primes the reference's
own benchmark uses (benchmark/bench_cryptography.cpp:24-36).

  value : pairs/s with pt and r already resident in HBM (CUDA events on the
          launching stream); under torchrun every rank has its own 65536-element
          shard (weak scaling), value = all units / max-over-ranks time
  e2e   : the same step through the host-pointer C ABI (ipclb200_encrypt +
          ipclb200_decrypt) from pinned host buffers, transfers inside the
          timing, over all `steps` (the library uses pinned operands in place:
          the kernels move them over PCIe while they compute; the same steps
          with staging copies are reported as e2e.staged_copies)
  e2e_ipcl : the same step through the C++ ipcl:: API an unmodified IPCL
          application calls (host BigNumbers in, host BigNumbers out; N = 1 only)
  roofline : the dominant kernel (two-digit CRT decrypt) against the measured
          IMAD.WIDE rate of this GPU (the path is integer-ALU bound, SURVEY.md
          section 8d): algorithmic fraction (generic w=5 modexp count) and the
          fraction of the multiplies it actually executes
  strong : ONE batch owned by rank 0 -- scatter -> encrypt -> decrypt -> gather
          inside the timing (NCCL send/recv) -- for the 2048-bit/65536 headline
          and BASELINE configs[3] (3072-bit key, 262144 elements); plus the same
          batch through the library's own one-process multi-GPU path
  configs : the other BASELINE configs (HE add / mul, raw modexp sweep, batch-8
          round trip), each with roofline fractions and a full-buffer oracle
          check of a sample
  cpu_baseline : the CPU restatement of the reference path (oracle/) on the
          host cores, bounded sample (N = 1 only)

`--impl reference` times the CPU path only (rank 0), see run_reference().
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KEY_BITS = 2048
BATCH = 65536
NL = KEY_BITS // 32          # words of n
RL = NL // 2                 # words of the DJN random (bits/2, pub_key.cpp:46)
PL = NL // 2                 # words of p, q

METRIC = "Paillier encrypt+decrypt ops/sec @2048-bit, batch 64K"
UNIT = "enc+dec pairs/s"


# algorithmic work per unit, SURVEY.md section 8(d): W(k,e) = N(e) * M(k/32)
def mont_macs(L):
    return 2 * L * L + L


def modexp_mults(e_bits):
    return e_bits + (e_bits + 4) // 5 + 32 + 2


def mac_encrypt(key_bits):
    return modexp_mults(key_bits // 2) * mont_macs(2 * key_bits // 32)


def mac_decrypt(key_bits):
    return 2 * modexp_mults(key_bits // 2) * mont_macs(key_bits // 32)


def mac_modexp(mod_bits, exp_bits):
    return modexp_mults(exp_bits) * mont_macs(mod_bits // 32)


def executed_products(exp_bits):
    """Montgomery products modexp_kernel executes for a per-element exponent:
    fixed window of the width pick_window() chooses (host_common.hpp), table
    build + squarings + window multiplies + into/out of Montgomery form"""
    best = None
    for w in range(1, 7):
        cost = ((1 << w) - 2) + exp_bits + (exp_bits + w - 1) // w
        best = cost if best is None else min(best, cost)
    return best + 2


MAC_ENCRYPT = mac_encrypt(KEY_BITS)      # 41.55 M
MAC_DECRYPT = mac_decrypt(KEY_BITS)      # 20.85 M
BYTES_ENCRYPT = NL * 4 + RL * 4 + 2 * NL * 4     # pt + r in, ct out
BYTES_DECRYPT = 2 * NL * 4 + NL * 4              # ct in, pt out


def hensel_modexp_macs(exp_bits, LH):
    """IMAD.WIDE of modexp_hensel_kernel per element (mod n^2, LH = words of n):
    enter 8 LH^2, table (2^w - 2) multiplies, exp_bits squarings, one multiply per
    window, leave 6 LH^2; squaring 4 LH^2 + LH, multiply 5 LH^2"""
    best_w, best = 1, None
    for w in range(1, 7):
        cost = ((1 << w) - 2) + exp_bits + (exp_bits + w - 1) // w
        if best is None or cost < best:
            best, best_w = cost, w
    w = best_w
    muls = ((1 << w) - 2) + (exp_bits + w - 1) // w
    return exp_bits * (4 * LH * LH + LH) + muls * 5 * LH * LH + 14 * LH * LH


def sliding_counts(e, w=5):
    """(squarings, multiplies) of the left-to-right sliding-window schedule the
    kernels run for a shared exponent (host_common.hpp: build_schedule)"""
    i, first, nsq, nmul = e.bit_length() - 1, True, 0, 0
    while i >= 0:
        if not (e >> i) & 1:
            nsq += 1
            i -= 1
            continue
        l = max(i - w + 1, 0)
        while not (e >> l) & 1:
            l += 1
        if not first:
            nsq += i - l + 1
            nmul += 1
        first = False
        i = l - 1
    return nsq, nmul


def hensel_window(p, q):
    """the sliding-window width ipclb200_privkey_create picks for the two-digit
    decrypt: fewest half-width products over both sides, 4..6 bits"""
    def cost(w):
        c = 0
        for pr in (p, q):
            nsq, nmul = sliding_counts(pr - 1, w)
            c += 4 * nsq + 5 * (nmul + (1 << (w - 1)) - 1)
        return c
    return min((4, 5, 6), key=cost)


def hensel_executed_macs(p, q):
    """IMAD.WIDE the two-digit decrypt executes per ciphertext (mont_hensel.cuh):
    a squaring is 4 LH^2 + LH, a multiply 5 LH^2, per side one squaring and
    2^(w-1) - 1 multiplies for the table of odd powers, 16 LH^2 to enter and
    2 LH^2 to leave"""
    total = 0
    w = hensel_window(p, q)
    for pr in (p, q):
        LH = (pr.bit_length() + 31) // 32
        nsq, nmul = sliding_counts(pr - 1, w)
        total += ((nsq + 1) * (4 * LH * LH + LH) + (nmul + (1 << (w - 1)) - 1) * 5 * LH * LH +
                  18 * LH * LH)
    return total


def comb_window(key_bits=KEY_BITS):
    """window width of the wide fixed-base table (comb_pick_window in
    csrc/ipcl_b200.cu): the widest <= 18 bits whose table of
    ceil(r_bits / w) * 2^w entries of 2n words fits the key's budget (4 GB)"""
    if "IPCLB200_COMB_WINDOW" in os.environ:
        return int(os.environ["IPCLB200_COMB_WINDOW"])
    budget = int(os.environ.get("IPCLB200_COMB_MAX_MB", "4096")) << 20
    r_bits, entry = key_bits // 2, 2 * key_bits // 8
    w = 18
    while w > 4 and ((r_bits + w - 1) // w) * (entry << w) > budget:
        w -= 1
    return w


def load_key(bits=KEY_BITS):
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[str(bits)].items()}
    p, q = sorted((k["p"], k["q"]))
    return p, q, k["hs"]


def synth_inputs(count, seed, nl=NL):
    """uniform plaintexts in [0, 2^(bits-2)) (< n) and uniform bits/2-bit randoms"""
    from pailliercryptolib_b200.limbs import random_limbs
    rng = np.random.default_rng(seed)
    pt = random_limbs(rng, count, nl, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, nl // 2)
    return pt, r


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU through NVML while the
    timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.err = repr(e)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "error": self.err}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def load_oracle(all_threads):
    """the CPU oracle (test infrastructure; here only as the timed CPU baseline
    and as the checker of the configs samples).  all_threads: torchrun exports
    OMP_NUM_THREADS=1 to its workers -- the reference arm must use the cores the
    box has, as ipcl's OMPUtilities::MaxThreads does (util.hpp:106-112)."""
    if all_threads and "oracle" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    return orc


def cpu_reference(count):
    """Times the CPU restatement of the reference path (oracle/) on `count`
    elements with all host threads: DJN encrypt then CRT decrypt, structured
    as the reference (generic fixed-window modexp per element; chunk-of-8
    multi-buffer AVX512-IFMA when the host has it).  Returns pairs/s etc."""
    orc = load_oracle(True)
    from pailliercryptolib_b200.limbs import to_limbs
    p, q, hs = load_key()
    n = p * q
    pt, r = synth_inputs(count, seed=0xB200)
    nl, hsl = to_limbs(n, NL), to_limbs(hs, 2 * NL)
    pl, ql = to_limbs(p, PL), to_limbs(q, PL)
    fast = getattr(orc, "have_ifma", lambda: False)()
    enc = orc.encrypt_mb8 if fast else orc.encrypt
    dec = orc.decrypt_crt_mb8 if fast else orc.decrypt_crt
    t0 = time.perf_counter()
    ct = enc(nl, hsl, pt, r)
    t1 = time.perf_counter()
    dt = dec(pl, ql, ct)
    t2 = time.perf_counter()
    assert np.array_equal(dt, pt), "CPU baseline round trip failed"
    return {
        "pairs_per_s": count / (t2 - t0),
        "encrypt_per_s": count / (t1 - t0),
        "decrypt_per_s": count / (t2 - t1),
        "seconds": t2 - t0,
        "cores": orc.num_threads(),
        "algo": ("8-lane AVX512-IFMA radix-2^52 fixed-window Montgomery "
                 "(restated mbx_exp_mb8) + OpenMP over chunks of 8" if fast else
                 "scalar radix-2^32 CIOS fixed-window Montgomery + OpenMP"),
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.
    The genuine library cannot be built here (IPP-Crypto is not vendored and
    there is no network), so this is the labelled port under
    oracle/.  Each step is the full 65536-element workload (--cpu-sample
    bounds it for quick checks), on every host thread of the box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample if args.cpu_sample > 0 else BATCH
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_reference(sample)
        if i >= args.warmup:
            times.append(res["seconds"])
    sec = float(np.mean(times))
    value = sample / sec
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (52-bit limbs)" if "IFMA" in res["algo"] else "u32",
        "data": "synthetic",
        "config": {"workload": "2048-bit key, batch=65536 encrypt+decrypt "
                               "(DJN r=1024 bit, CRT decrypt)",
                   "elements_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"],
                         "kind": "port",
                         "sample": "%d elements per step, %s" % (sample, res["algo"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def ipcl_e2e(batch, steps):
    """The headline workload through the C++ ipcl:: API (vector<BigNumber> in,
    vector<BigNumber> out, benchmarks/bench_ipcl.cpp --e2e) in a child process:
    once with device-resident texts (the library's default: the ciphertexts stay
    in HBM between encrypt and decrypt) and once with a host round trip of every
    text (IPCL_B200_DEVICE_RESIDENT=0)."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "bench_ipcl")
    if not os.path.exists(exe):
        return {"unavailable": "tests/cpp/_build/bench_ipcl not built"}
    out = {"unit": UNIT, "batch": batch, "steps": steps,
           "api": "ipcl::PublicKey::encrypt -> ipcl::PrivateKey::decrypt -> getTexts(), "
                  "wall clock in the C++ caller, PlainText built from host BigNumbers every "
                  "step, all plaintexts compared with their inputs"}
    for key, resident in (("resident_texts", "1"), ("host_round_trip", "0")):
        env = dict(os.environ, IPCL_B200_DEVICE_RESIDENT=resident, IPCLB200_COMB_SYNC="1")
        env.pop("OMP_NUM_THREADS", None)
        try:
            r = subprocess.run([exe, "--e2e", str(batch), str(steps)], env=env, text=True,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out[key] = {"value": d["pairs_per_s"], "ms_per_step": d["ms_per_step"],
                        "mismatches": d["mismatches"]}
        except Exception as e:  # reported, never fatal for the headline
            out[key] = {"error": repr(e)[:200]}
    if "value" in out.get("resident_texts", {}):
        out["value"] = out["resident_texts"]["value"]
    return out


def emit_json(line):
    """stdout carries exactly ONE line: the JSON.  Libraries that write to the
    C-level stdout (NCCL prints its version banner there) were redirected to
    stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


# ---------------------------------------------------------------------------
# strong scaling: one batch owned by rank 0
# ---------------------------------------------------------------------------
def strong_case(torch, dist, capi, sharding, dev, stream, world, rank, bits, total, steps):
    """ONE `total`-element batch resident on rank 0's GPU: scatter plaintexts and
    randoms -> encrypt -> decrypt -> gather the plaintexts back, all inside the
    timed region.  Returns (ms per step max over ranks, scatter+gather ms)."""
    from pailliercryptolib_b200.limbs import to_limbs
    nl = bits // 32
    p, q, hs = load_key(bits)
    pk = capi.PubKey(to_limbs(p * q, nl), to_limbs(hs, 2 * nl), bits // 2)
    sk = capi.PrivKey(to_limbs(p, nl // 2), to_limbs(q, nl // 2))
    full_pt = full_r = None
    if rank == 0:
        pt_h, r_h = synth_inputs(total, seed=0x5712 + bits, nl=nl)
        full_pt = torch.from_numpy(pt_h.view(np.int32)).to(dev)
        full_r = torch.from_numpy(r_h.view(np.int32)).to(dev)
    s, e = sharding.shard_range(total, world, rank)
    cnt = e - s
    d_ct = torch.empty((cnt, 2 * nl), dtype=torch.int32, device=dev)
    d_dt = torch.empty((cnt, nl), dtype=torch.int32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(ev=None):
        if ev:
            ev[0].record()
        loc_pt = sharding.scatter_rows(full_pt, total, nl, torch.int32, dev)
        loc_r = sharding.scatter_rows(full_r, total, nl // 2, torch.int32, dev)
        if ev:
            ev[1].record()
        pk.encrypt_dev(loc_pt.data_ptr(), nl, loc_r.data_ptr(), nl // 2, cnt,
                       d_ct.data_ptr(), stream)
        sk.decrypt_dev(d_ct.data_ptr(), cnt, d_dt.data_ptr(), stream)
        if ev:
            ev[2].record()
        out = sharding.gather_rows(d_dt, total)
        if ev:
            ev[3].record()
        return out

    os.environ["IPCLB200_COMB_SYNC"] = "1"   # wide table before the timing starts
    for _ in range(2):
        out = one()
    barrier()
    if rank == 0:
        assert torch.equal(out, full_pt), "strong: decrypt(encrypt(pt)) != pt"
    ms, sg = [], []
    for _ in range(steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        one(ev)
        ev[3].synchronize()
        ms.append(ev[0].elapsed_time(ev[3]))
        sg.append(ev[0].elapsed_time(ev[1]) + ev[2].elapsed_time(ev[3]))
    barrier()
    step_ms, sg_ms = sharding.max_over_ranks([float(np.mean(ms)), float(np.mean(sg))], dev)
    del pk, sk
    return step_ms, sg_ms


def inproc_case(torch, capi, ndev, bits, total, steps):
    """the same batch through the library's own multi-GPU path: ONE process,
    ipclb200_init_devices(ndev), host-pointer encrypt + decrypt from pinned host
    memory (every device copies its own contiguous block)"""
    from pailliercryptolib_b200.limbs import to_limbs
    nl = bits // 32
    got = capi.init_devices(ndev)
    p, q, hs = load_key(bits)
    pk = capi.PubKey(to_limbs(p * q, nl), to_limbs(hs, 2 * nl), bits // 2)
    sk = capi.PrivKey(to_limbs(p, nl // 2), to_limbs(q, nl // 2))
    pt_h, r_h = synth_inputs(total, seed=0x1A + bits, nl=nl)
    pin = lambda a: torch.from_numpy(a.view(np.int32)).pin_memory()
    pt_pin, r_pin = pin(pt_h), pin(r_h)
    ct_pin = torch.empty((total, 2 * nl), dtype=torch.int32).pin_memory()
    dt_pin = torch.empty((total, nl), dtype=torch.int32).pin_memory()
    u32 = lambda t: t.numpy().view(np.uint32)

    def one():
        pk.encrypt(u32(pt_pin), u32(r_pin), out=u32(ct_pin))
        sk.decrypt(u32(ct_pin), out=u32(dt_pin))

    os.environ["IPCLB200_COMB_SYNC"] = "1"
    for _ in range(2):
        one()
    assert np.array_equal(u32(dt_pin), pt_h), "inproc: round trip failed"
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    del pk, sk
    return got, ms


# ---------------------------------------------------------------------------
# the other BASELINE configs (rank 0, devices of the library's active set)
# ---------------------------------------------------------------------------
def run_configs(torch, capi, peak_mac, quick):
    from pailliercryptolib_b200.limbs import random_limbs, to_limbs
    orc = load_oracle(True)
    fast = orc.have_ifma()
    out = []
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0xC0F)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    def cuda(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)

    def host(t):
        return t.cpu().numpy().view(np.uint32)

    # ---- configs[2]: HE add (ct+ct) and HE mul (ct*pt), 2048-bit key, 65536 ----
    p, q, hs = load_key()
    n = p * q
    nsq = to_limbs(n * n, 2 * NL)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(hs, 2 * NL), KEY_BITS // 2)
    count = 16384 if quick else BATCH
    pt, r = synth_inputs(count, 0xADD)
    d_a = torch.empty((count, 2 * NL), dtype=torch.int32, device=dev)
    pk.encrypt_dev(cuda(pt).data_ptr(), NL, cuda(r).data_ptr(), RL, count, d_a.data_ptr(),
                   stream)
    torch.cuda.synchronize()
    d_b = d_a.flip(0).contiguous()
    d_o = torch.empty_like(d_a)
    ms = timed(lambda: capi.modmul_dev(d_a.data_ptr(), d_b.data_ptr(), nsq, count,
                                       d_o.data_ptr(), stream), 5)
    S = 2048
    ok = bool(np.array_equal(host(d_o[:S]), orc.modmul(host(d_a[:S]), host(d_b[:S]), nsq)))
    macs = 2 * mont_macs(2 * NL)
    out.append({"config": "2048-bit key, batch=%d HE add (ct+ct)" % count, "ms": ms,
                "ops_per_s": count / ms * 1e3, "kernel": "modmul_kernel<16,8>",
                "roofline": {"frac": count * macs / (ms * 1e-3) / peak_mac,
                             "executed_frac": count * 2 * 2 * (2 * NL) ** 2 / (ms * 1e-3) / peak_mac,
                             "hbm_gbs": count * 3 * 2 * NL * 4 / (ms * 1e-3) / 1e9},
                "oracle_check": {"elements": S, "ok": ok}})
    for ebits, label in ((32, "32-bit plaintexts"), (2048, "2048-bit plaintexts")):
        ew = max(1, ebits // 32)
        cnt = count if ebits == 32 else (4096 if quick else BATCH)
        e = random_limbs(rng, cnt, ew)
        d_e = cuda(e)
        ms = timed(lambda: capi.modexp_dev(d_a.data_ptr(), d_e.data_ptr(), nsq, ew, ebits, cnt,
                                           d_o.data_ptr(), stream), 2 if ebits == 32 else 1)
        S2 = 1024
        want = (orc.modexp_mb8(host(d_a[:S2]), e[:S2], nsq) if fast else
                orc.modexp(host(d_a[:S2]), e[:S2], nsq[None, :], shared_mod=True))
        ok = bool(np.array_equal(host(d_o[:S2]), want))
        macs = mac_modexp(2 * KEY_BITS, ebits)
        out.append({"config": "2048-bit key, batch=%d HE mul (ct*pt), %s" % (cnt, label),
                    "ms": ms, "ops_per_s": cnt / ms * 1e3,
                    "kernel": "modexp_hensel_kernel<16,4> (two-digit arithmetic mod n^2)",
                    "roofline": {"frac": cnt * macs / (ms * 1e-3) / peak_mac,
                                 "executed_frac": cnt * hensel_modexp_macs(ebits, NL) /
                                 (ms * 1e-3) / peak_mac,
                                 "note": "frac counts the generic full-width algorithm "
                                         "(SURVEY 8d), executed_frac the multiplies run"},
                    "oracle_check": {"elements": S2, "ok": ok}})
    del pk
    # ---- configs[4]: raw modexp sweep (k-bit modulus, k-bit exponent) -----------
    for k in (1024, 2048, 3072, 4096):
        L = k // 32
        mod = random_limbs(rng, 1, L)
        mod[0, 0] |= 1
        mod[0, -1] |= 0x80000000
        for lg in ((10, 14) if quick else (10, 14, 18)):
            cnt = 1 << lg
            base = random_limbs(rng, cnt, L)
            base[:, -1] &= 0x7FFFFFFF
            e = random_limbs(rng, cnt, L)
            d_base, d_e = cuda(base), cuda(e)
            d_out = torch.empty_like(d_base)
            reps = 1 if (lg == 18 and k >= 3072) else 2
            ms = timed(lambda: capi.modexp_dev(d_base.data_ptr(), d_e.data_ptr(), mod[0], L, k,
                                               cnt, d_out.data_ptr(), stream), reps)
            S3 = min(cnt, 1024 if k <= 2048 else 256)
            want = (orc.modexp_mb8(base[:S3], e[:S3], mod[0]) if fast and k <= 4096 else
                    orc.modexp(base[:S3], e[:S3], mod, shared_mod=True))
            ok = bool(np.array_equal(host(d_out[:S3]), want))
            macs = mac_modexp(k, k)
            nprod = executed_products(k)
            out.append({"config": "raw modexp %d-bit, batch=2^%d" % (k, lg), "ms": ms,
                        "ops_per_s": cnt / ms * 1e3,
                        "roofline": {"frac": cnt * macs / (ms * 1e-3) / peak_mac,
                                     "executed_frac": cnt * nprod * 2 * L * L / (ms * 1e-3) / peak_mac,
                                     "hbm_gbs": cnt * 3 * L * 4 / (ms * 1e-3) / 1e9},
                        "oracle_check": {"elements": S3, "ok": ok}})
            del d_base, d_e, d_out
    # ---- configs[0]: 1024-bit key, batch 8 round trip (host-pointer C ABI) ------
    p1, q1, hs1 = load_key(1024)
    n1 = p1 * q1
    pk1 = capi.PubKey(to_limbs(n1, 32), to_limbs(hs1, 64), 512)
    sk1 = capi.PrivKey(to_limbs(p1, 16), to_limbs(q1, 16))
    pt8 = random_limbs(rng, 8, 32, top_mask=0x3FFFFFFF)
    r8 = random_limbs(rng, 8, 16)
    ct8 = pk1.encrypt(pt8, r8)
    t0 = time.perf_counter()
    for _ in range(20):
        ct8 = pk1.encrypt(pt8, r8)
        dt8 = sk1.decrypt(ct8)
    ms = (time.perf_counter() - t0) * 1e3 / 20
    ok = bool(np.array_equal(ct8, orc.encrypt(to_limbs(n1, 32), to_limbs(hs1, 64), pt8, r8))
              and np.array_equal(dt8, pt8))
    out.append({"config": "1024-bit key, batch=8 encrypt->decrypt round trip "
                          "(host-pointer C ABI, wall clock)", "ms": ms,
                "ops_per_s": 8 / ms * 1e3, "oracle_check": {"elements": 8, "ok": ok}})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="elements of the workload timed on the CPU (0 = 65536 for "
                         "--impl reference, 16384 for the cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-ipcl", action="store_true", help="skip the e2e_ipcl leg")
    ap.add_argument("--quick", action="store_true", help="smaller side measurements")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from pailliercryptolib_b200 import capi, sharding
    from pailliercryptolib_b200.limbs import to_limbs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side barrier for the phases in which rank 0 alone drives every GPU
        # of the box: a NCCL barrier would keep a spinning kernel on the other
        # ranks' GPUs and time-slice with rank 0's work there
        cpu_group = dist.new_group(backend="gloo")
    capi.init(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream

    p, q, hs = load_key()
    n = p * q
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(hs, 2 * NL), KEY_BITS // 2)
    sk = capi.PrivKey(to_limbs(p, PL), to_limbs(q, PL))

    B = args.batch
    pt_h, r_h = synth_inputs(B, seed=0xB200 + rank)
    # pinned host buffers for the end-to-end leg
    pin = lambda a: torch.from_numpy(a.view(np.int32)).pin_memory()
    pt_pin, r_pin = pin(pt_h), pin(r_h)
    ct_pin = torch.empty((B, 2 * NL), dtype=torch.int32).pin_memory()
    dt_pin = torch.empty((B, NL), dtype=torch.int32).pin_memory()
    # device-resident inputs for the kernel leg
    d_pt = pt_pin.to(dev)
    d_r = r_pin.to(dev)
    d_ct = torch.empty((B, 2 * NL), dtype=torch.int32, device=dev)
    d_dt = torch.empty((B, NL), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)

    def step_dev(ev=None):
        if ev:
            ev[0].record()
        pk.encrypt_dev(d_pt.data_ptr(), NL, d_r.data_ptr(), RL, B,
                       d_ct.data_ptr(), stream)
        if ev:
            ev[1].record()
        sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), stream)
        if ev:
            ev[2].record()

    def step_e2e():
        pk.encrypt(pt_pin.numpy().view(np.uint32), r_pin.numpy().view(np.uint32),
                   out=ct_pin.numpy().view(np.uint32))
        sk.decrypt(ct_pin.numpy().view(np.uint32), out=dt_pin.numpy().view(np.uint32))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- integer roofline denominator, measured live on this GPU -----------
    peak_mac, _ = capi.int_peak()

    # ---- warm-up: also builds the fixed-base table.  The wide (16-bit window,
    # 2.1 GB) table is normally built on a side stream while the first calls use
    # the 17 MB one; here the bench waits for it so that every timed step runs on
    # the steady-state table.  Its build time is reported below.
    os.environ["IPCLB200_COMB_SYNC"] = "1"
    t_tab0 = time.perf_counter()
    step_dev()
    torch.cuda.synchronize()
    table_build_ms = (time.perf_counter() - t_tab0) * 1e3
    for _ in range(max(args.warmup, 3) - 1):
        step_dev()
    torch.cuda.synchronize()
    assert torch.equal(d_dt, d_pt), "decrypt(encrypt(pt)) != pt on the device"

    # ---- timed: device-resident -------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = capi.launch_count()
    enc_ms, dec_ms = [], []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        step_dev(ev)
        ev[2].synchronize()
        enc_ms.append(ev[0].elapsed_time(ev[1]))
        dec_ms.append(ev[1].elapsed_time(ev[2]))
        flush.fill_(1)   # > L2 (126 MB) written between timed iterations
    barrier()
    t_wall1 = time.perf_counter()
    launches = capi.launch_count() - launches0
    step_ms = float(np.sum(enc_ms) + np.sum(dec_ms)) / args.steps
    # ---- timed: end to end through the host-pointer C ABI, all steps ---------
    # The pinned caller buffers are used in place by the kernels (zero-copy over
    # PCIe, include/ipcl_b200.h: ipclb200_zero_copy_count); the same steps with
    # staging copies before / after every launch are timed next to it.
    def time_e2e():
        step_e2e()   # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        return (time.perf_counter() - t0) * 1e3 / args.steps

    os.environ["IPCLB200_ZERO_COPY"] = "0"
    e2e_staged_ms = time_e2e()
    del os.environ["IPCLB200_ZERO_COPY"]
    zc0 = capi.zero_copy_count()
    e2e_ms = time_e2e()
    in_place = (capi.zero_copy_count() - zc0) / (args.steps + 1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert np.array_equal(dt_pin.numpy().view(np.uint32), pt_h), "e2e round trip failed"

    step_ms, e2e_ms, enc_mean, dec_mean, e2e_staged_ms = sharding.max_over_ranks(
        [step_ms, e2e_ms, float(np.mean(enc_ms)), float(np.mean(dec_ms)), e2e_staged_ms], dev)
    del d_ct, d_dt, flush

    # ---- strong scaling: one batch owned by rank 0 ---------------------------
    strong = None
    if not args.no_strong:
        ssteps = max(1, min(args.steps, 5))
        strong = {"note": "ONE batch resident on rank 0's GPU; scatter (NCCL send/recv) -> "
                          "encrypt -> decrypt -> gather inside the timing, max over ranks"}
        ms, sg = strong_case(torch, dist, capi, sharding, dev, stream, world, rank,
                             KEY_BITS, BATCH, ssteps)
        strong["headline_64k"] = {
            "workload": "2048-bit key, ONE batch of 65536", "ms_per_step": ms,
            "value": BATCH / (ms * 1e-3), "unit": UNIT, "scatter_gather_ms": sg,
            "steps": ssteps}
        c4 = 65536 if args.quick else 262144
        ms, sg = strong_case(torch, dist, capi, sharding, dev, stream, world, rank,
                             3072, c4, max(1, min(ssteps, 3)))
        strong["config4_3072bit"] = {
            "workload": "3072-bit key, ONE batch of %d (BASELINE configs[3])" % c4,
            "ms_per_step": ms, "value": c4 / (ms * 1e-3), "unit": UNIT,
            "scatter_gather_ms": sg, "steps": max(1, min(ssteps, 3))}
    barrier()
    if rank == 0 and not args.no_strong:
        # the library's own one-process path over the same GPUs (the other ranks
        # wait in a CPU barrier, their GPUs are idle)
        ndev = min(world, torch.cuda.device_count())
        got, ms = inproc_case(torch, capi, ndev, KEY_BITS, BATCH, max(1, min(args.steps, 5)))
        strong["inproc_64k"] = {
            "workload": "2048-bit key, ONE batch of 65536 in pinned host memory, ONE "
                        "process driving %d GPU(s) through ipclb200_init_devices + the "
                        "host-pointer C ABI (copies inside the timing)" % got,
            "devices": got, "ms_per_step": ms, "value": BATCH / (ms * 1e-3), "unit": UNIT}
        capi.init_devices(1) if local == 0 else None
    e2e_ipcl = None
    if rank == 0 and world == 1 and not args.no_ipcl:
        e2e_ipcl = ipcl_e2e(BATCH, max(1, min(args.steps, 5)))
    configs = None
    if rank == 0 and not args.no_configs:
        torch.cuda.set_device(local)
        configs = run_configs(torch, capi, peak_mac, args.quick or world > 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(group=cpu_group)

    if rank == 0:
        total = B * world
        value = total / (step_ms * 1e-3)
        dec_s = dec_mean * 1e-3
        enc_s = enc_mean * 1e-3
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f)["hbm_gbs"]
            hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass
        comb_w = comb_window()
        comb_windows = (KEY_BITS // 2 + comb_w - 1) // comb_w
        enc_exec = ((comb_windows - 1) * 5 + 5) * NL * NL
        ach = B * MAC_DECRYPT / dec_s / 1e12
        exec_macs = hensel_executed_macs(p, q)
        roofline = {
            "kernel": "decrypt_hensel_kernel<32,1,3,4,W64,128,compact> + crt_combine_kernel "
                      "(per ciphertext two 1024-bit-exponent modexps mod p^2, q^2 in two-digit "
                      "arithmetic, one (ciphertext, side) task per thread, 12 warps/SM)",
            "bound": "int32_alu",
            "achieved": ach, "peak": peak_mac / 1e12, "unit": "TMAC32/s",
            "frac": ach / (peak_mac / 1e12),
            "executed_frac": B * exec_macs / dec_s / peak_mac,
            "note": "frac counts the generic algorithm of SURVEY 8d (full-width w=5 "
                    "windowed modexp, 20.85 M MAC32 per ciphertext) and exceeds 1 because "
                    "the kernel works in half-width digits mod p^2 (%.2f M MAC32 executed); "
                    "executed_frac is the share of the IMAD.WIDE issue rate the executed "
                    "multiplies take" % (exec_macs / 1e6),
            "peak_source": "measured live: dependent IMAD.WIDE.U32 chains on all "
                           "SMs (ipclb200_int_peak)",
            "algorithmic_mac32_per_unit": MAC_DECRYPT,
            "executed_mac32_per_unit": exec_macs,
            "units_per_launch": B,
            "launch_ms": dec_mean,
            "traffic": (traffic or {}).get("dram_bytes_per_launch"),
            "traffic_source": (traffic or {}).get(
                "source", "none") + " (static: read from profiles/ncu_traffic.json, "
                                    "not measured in this run)",
            "traffic_note": (traffic or {}).get("note"),
            "hbm": {"achieved": B * BYTES_DECRYPT / dec_s / 1e9, "peak": hbm_peak,
                    "unit": "GB/s",
                    "frac": B * BYTES_DECRYPT / dec_s / 1e9 / hbm_peak,
                    "peak_source": hbm_src},
            "encrypt_kernel": {
                "kernel": "encrypt_hensel_kernel<16,4> (fixed-base table of pairs for hs^r, "
                          "two-digit arithmetic mod n^2)",
                "achieved": B * MAC_ENCRYPT / enc_s / 1e12, "unit": "TMAC32/s",
                "frac": B * MAC_ENCRYPT / enc_s / peak_mac,
                "note": "algorithmic count is the generic w=5 windowed modexp mod n^2 "
                        "(41.55 M MAC32); the kernel multiplies one table entry per %d-bit "
                        "window (%d two-digit products of 5*64^2 MAC32) and finishes with "
                        "a*m mod n and t*n (5*64^2): %.2f M MAC32 executed, so frac > 1 is "
                        "the algorithm, not the pipe" % (comb_w, comb_windows - 1,
                                                         enc_exec / 1e6),
                "executed_frac": B * enc_exec / enc_s / peak_mac,
                "launch_ms": enc_mean,
            },
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (32-bit limbs, 64-bit multiply-add)",
            "data": "synthetic",
            "config": {"workload": "2048-bit key, batch=65536 encrypt+decrypt "
                                   "(DJN r=1024 bit, CRT decrypt), 1xB200 per rank",
                       "batch_per_gpu": B, "key_bits": KEY_BITS,
                       "l2": "256 MB flush written between timed iterations",
                       "parallelism": "shard per GPU, no data-path collective",
                       "fixed_base_table": "%d-bit windows, %.1f GB, built once per key in "
                                           "warm-up (first step incl. build: %.0f ms)"
                                           % (comb_window(),
                                              ((KEY_BITS // 2 + comb_window() - 1) //
                                               comb_window()) * (512 << comb_window()) / 2 ** 30,
                                              table_build_ms)},
            "encrypt_per_s": total / enc_s, "decrypt_per_s": total / dec_s,
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms, "steps": args.steps,
                    "h2d_bytes_per_step": B * (NL + RL + 2 * NL) * 4,
                    "d2h_bytes_per_step": B * (2 * NL + NL) * 4,
                    "in_place_operands_per_step": in_place,
                    "transfer": "encrypt reads its plaintexts and writes its ciphertexts "
                                "through the mapped pinned buffers while it runs; randoms, "
                                "the ciphertexts of decrypt and the decrypted plaintexts go "
                                "by copy engines",
                    "staged_copies": {"value": total / (e2e_staged_ms * 1e-3),
                                      "ms_per_step": e2e_staged_ms,
                                      "note": "IPCLB200_ZERO_COPY=0: every operand staged "
                                              "by a copy before / after the launch"}},
            "gpu_launches": launches,
            "roofline": roofline,
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall1 - t_wall0,
        }
        if e2e_ipcl is not None:
            line["e2e_ipcl"] = e2e_ipcl
        if strong is not None:
            line["strong"] = strong
        if configs is not None:
            line["configs"] = configs
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample if args.cpu_sample > 0 else 16384
            res = cpu_reference(sample)
            line["cpu_baseline"] = {
                "value": res["pairs_per_s"], "unit": UNIT, "cores": res["cores"],
                "kind": "port",
                "sample": "%d of 65536 elements, %s" % (sample, res["algo"]),
                "encrypt_per_s": res["encrypt_per_s"],
                "decrypt_per_s": res["decrypt_per_s"]}
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
