#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json:
Paillier encrypt+decrypt throughput at a 2048-bit key on a 65536-element batch.

One "step" = encrypt 65536 plaintexts (DJN obfuscator hs^r, r = 1024 bit, the
default key type of ipcl::generateKeypair) and CRT-decrypt the 65536
ciphertexts that came out.  Key = the ISO/IEC 18033-6 primes the reference's
own benchmark uses (benchmark/bench_cryptography.cpp:24-36).

  value : pairs/s with pt and r already resident in HBM (CUDA events on the
          launching stream)
  e2e   : the same step through the host-pointer C ABI (ipclb200_encrypt +
          ipclb200_decrypt) from pinned host buffers, copies inside the timing
  roofline : the decrypt modexp kernel against the measured IMAD.WIDE rate of
          this GPU (the path is integer-ALU bound), plus
          its HBM figures
  cpu_baseline : the CPU restatement of the reference path (oracle/) on the
          host cores, bounded sample

Multi-GPU: one process per GPU (torchrun), every rank encrypts+decrypts its own
65536-element shard -- independent units, no data-path collective -- weak
scaling; value = all units / max-over-ranks time.

`--impl reference` times the CPU path only (rank 0), see cpu_reference().
"""
import argparse
import json
import os
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


MAC_ENCRYPT = modexp_mults(KEY_BITS // 2) * mont_macs(2 * NL)      # 41.55 M
MAC_DECRYPT = 2 * modexp_mults(KEY_BITS // 2) * mont_macs(NL)      # 20.85 M
BYTES_ENCRYPT = NL * 4 + RL * 4 + 2 * NL * 4     # pt + r in, ct out
BYTES_DECRYPT = 2 * NL * 4 + NL * 4              # ct in, pt out


def load_key():
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[str(KEY_BITS)].items()}
    p, q = sorted((k["p"], k["q"]))
    return p, q, k["hs"]


def synth_inputs(count, seed):
    """uniform plaintexts in [0, 2^2046) (< n) and uniform 1024-bit randoms"""
    from pailliercryptolib_b200.limbs import random_limbs
    rng = np.random.default_rng(seed)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, RL)
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


def cpu_reference(count, threads_note=True):
    """Times the CPU restatement of the reference path (oracle/) on `count`
    elements with all host threads: DJN encrypt then CRT decrypt, structured
    as the reference (generic fixed-window modexp per element; chunk-of-8
    multi-buffer AVX512-IFMA when the host has it).  Returns pairs/s etc."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
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
    oracle/.  Each step is a bounded sample of the 65536-element workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
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
        "ms_per_step": sec * 1e3 * BATCH / sample,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (52-bit limbs)" if "IFMA" in res["algo"] else "u32",
        "data": "synthetic",
        "config": {"workload": "2048-bit key, batch=65536 encrypt+decrypt "
                               "(DJN r=1024 bit, CRT decrypt)",
                   "sample": "%d of 65536 elements per step" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"],
                         "kind": "port",
                         "sample": "%d elements, %s" % (sample, res["algo"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def emit_json(line):
    """stdout carries exactly ONE line: the JSON.  Libraries that write to the
    C-level stdout (NCCL prints its version banner there) were redirected to
    stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-sample", type=int, default=16384,
                    help="elements of the workload timed on the CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from pailliercryptolib_b200 import capi
    from pailliercryptolib_b200.limbs import to_limbs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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

    # ---- warm-up (also builds the comb table and sizes the workspaces) -----
    for _ in range(max(args.warmup, 3)):
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
    # ---- timed: end to end through the host-pointer C ABI ------------------
    step_e2e()   # warm
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert np.array_equal(dt_pin.numpy().view(np.uint32), pt_h), "e2e round trip failed"

    from pailliercryptolib_b200 import sharding
    # the scatter/gather a single-owner batch would need (north_star: "NCCL
    # only for the trivial scatter/gather"); outside the timed region
    sg_ms = None
    if world > 1:
        full = d_pt.repeat(world, 1) if rank == 0 else None
        for _ in range(2):
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            barrier()
            e0.record()
            loc = sharding.scatter_rows(full, B * world, NL, torch.int32, dev)
            back = sharding.gather_rows(loc, B * world)
            e1.record()
            e1.synchronize()
            sg_ms = e0.elapsed_time(e1)
        del full, loc, back
    step_ms, e2e_ms, enc_mean, dec_mean = sharding.max_over_ranks(
        [step_ms, e2e_ms, float(np.mean(enc_ms)), float(np.mean(dec_ms))], dev)
    if sg_ms is not None:
        sg_ms = sharding.max_over_ranks([sg_ms], dev)[0]

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
                traffic = json.load(f).get("decrypt_crt_kernel_dram_bytes")
        except Exception:
            pass
        comb_w = int(os.environ.get("IPCLB200_COMB_WINDOW", "16"))
        comb_products = (KEY_BITS // 2 + comb_w - 1) // comb_w - 1 + 3
        ach = B * MAC_DECRYPT / dec_s / 1e12
        roofline = {
            "kernel": "decrypt_crt_kernel<16,4> (2 x 2048-bit-modulus, "
                      "1024-bit-exponent modexp per ciphertext)",
            "bound": "int32_alu",
            "achieved": ach, "peak": peak_mac / 1e12, "unit": "TMAC32/s",
            "frac": ach / (peak_mac / 1e12),
            "peak_source": "measured live: dependent IMAD.WIDE.U32 chains on all "
                           "SMs (ipclb200_int_peak)",
            "algorithmic_mac32_per_unit": MAC_DECRYPT,
            "units_per_launch": B,
            "launch_ms": dec_mean,
            "traffic": traffic,
            "hbm": {"achieved": B * BYTES_DECRYPT / dec_s / 1e9, "peak": hbm_peak,
                    "unit": "GB/s",
                    "frac": B * BYTES_DECRYPT / dec_s / 1e9 / hbm_peak,
                    "peak_source": hbm_src},
            "encrypt_kernel": {
                "kernel": "encrypt_kernel<16,8> (fixed-base comb for hs^r)",
                "achieved": B * MAC_ENCRYPT / enc_s / 1e12, "unit": "TMAC32/s",
                "frac": B * MAC_ENCRYPT / enc_s / peak_mac,
                "note": "algorithmic count is the generic w=5 windowed modexp "
                        "(41.55 M MAC32); the comb kernel (%d-bit windows) executes "
                        "%d Montgomery products (%.2f M MAC32), so frac > 1 is the "
                        "algorithm, not the pipe" % (comb_w, comb_products,
                                                     comb_products * 2 * (2 * NL) ** 2 / 1e6),
                "executed_frac": B * comb_products * 2 * (2 * NL) ** 2 / enc_s / peak_mac,
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
                       "parallelism": "shard per GPU, no data-path collective"},
            "encrypt_per_s": total / enc_s, "decrypt_per_s": total / dec_s,
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": B * (NL + RL + 2 * NL) * 4,
                    "d2h_bytes_per_step": B * (2 * NL + NL) * 4},
            "gpu_launches": launches,
            "roofline": roofline,
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall1 - t_wall0,
        }
        if sg_ms is not None:
            line["nccl_scatter_gather_ms"] = sg_ms
        if not args.no_cpu_baseline and world == 1:
            res = cpu_reference(args.cpu_sample)
            line["cpu_baseline"] = {
                "value": res["pairs_per_s"], "unit": UNIT, "cores": res["cores"],
                "kind": "port",
                "sample": "%d of 65536 elements, %s" % (args.cpu_sample, res["algo"]),
                "encrypt_per_s": res["encrypt_per_s"],
                "decrypt_per_s": res["decrypt_per_s"]}
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
