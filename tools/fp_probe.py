"""GPU probe for the FP64-pipe / dual-pipe CRT decrypt kernels.

1. stage-by-stage check of the FP64 role against exact integer arithmetic
   (IPCLB200_FP_DEBUG_STAGE), so that a wrong result is localised in one run;
2. residues of every pipe configuration against Python pow();
3. device-resident timing of every configuration at the bench batch.

Run on a B200:  python tools/fp_probe.py [--count 65536] [--skip-stages]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

from model_fp64_mont import build_schedule  # noqa: E402

W, FL = 22, 96
R = 1 << (W * FL)


def mont(a, b, n, npr):
    t = a * b
    q = (t * npr) % R
    return (t + q * n) // R


def expected_stage(ct, prime, stage):
    n = prime * prime
    npr = (-pow(n, -1, R)) % R
    v = mont(ct % R, 1, n, npr)
    if stage == 1:
        return v
    v += ct >> (W * FL)
    if stage == 2:
        return v
    x = mont(v, pow(R, 3, n), n, npr)
    if stage == 3:
        return x
    x2 = mont(x, x, n, npr)
    if stage == 4:
        return x2
    sched = build_schedule(prime - 1)
    tab = [x]
    t = x
    for _ in range(1, sched[0]):
        t = mont(t, x2, n, npr)
        tab.append(t)
    if stage == 5:
        return t
    acc = tab[sched[1]]
    for op in sched[2:]:
        if op == 0xFF:
            break
        acc = mont(acc if op == 0 else tab[op - 1], acc, n, npr)
    if stage == 6:
        return acc
    r = mont(acc, 1, n, npr)
    return 0 if r == n else r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=65536)
    ap.add_argument("--small", type=int, default=203)
    ap.add_argument("--skip-stages", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="")
    args = ap.parse_args()

    import torch
    from pailliercryptolib_b200 import capi
    from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,
                                              random_limbs, to_limbs)
    capi.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)["2048"].items()}
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    nsq = n * n
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    pk = capi.PubKey(to_limbs(n, 64), to_limbs(k["hs"], 128), 1024)
    rng = np.random.default_rng(7)
    cts = [int.from_bytes(rng.bytes(512), "little") % nsq for _ in range(args.small)]
    cts[0], cts[1], cts[2], cts[3] = 0, 1, nsq - 1, p * 12345
    cts[4] = q * q * 3 % nsq
    ct_small = batch_to_limbs(cts, 128)
    ok_all = True

    def set_mode(mode, **env):
        for key in ("IPCLB200_DECRYPT", "IPCLB200_FP_BLOCKS", "IPCLB200_FP_MASK",
                    "IPCLB200_DUAL2", "IPCLB200_FP_DEBUG_STAGE"):
            os.environ.pop(key, None)
        os.environ["IPCLB200_DECRYPT"] = mode
        for a, b in env.items():
            os.environ[a] = str(b)

    if not args.skip_stages:
        nchk = 24
        for stage in (1, 2, 3, 4, 5, 6, 0):
            set_mode("fp", IPCLB200_FP_DEBUG_STAGE=stage)
            x = sk.crt_residues(ct_small[:nchk])
            bad = []
            for i in range(nchk):
                for side, prime in ((0, p), (1, q)):
                    got = int.from_bytes(x[i, side].tobytes(), "little")
                    want = expected_stage(cts[i], prime, stage) % (1 << 2048)
                    if got != want:
                        bad.append((i, side, got, want))
            print("stage %d: %s (%d of %d wrong)" % (
                stage, "ok" if not bad else "MISMATCH", len(bad), 2 * nchk), flush=True)
            if bad:
                ok_all = False
                i, side, got, want = bad[0]
                print("  first: element %d side %d" % (i, side))
                print("  got  %x" % got)
                print("  want %x" % want)
                d = got ^ want
                print("  differing bits: lowest %d highest %d" % (
                    (d & -d).bit_length() - 1, d.bit_length() - 1))
                print("  wrong (element, side): %s" % [(b[0], b[1]) for b in bad[:48]])
                break

    want_res = None
    configs = [
        ("int", {}),
        ("fp", {"IPCLB200_FP_BLOCKS": 2}),
        ("fp", {"IPCLB200_FP_BLOCKS": 3}),
        ("fp", {"IPCLB200_FP_BLOCKS": 1}),
        ("dual", {"IPCLB200_FP_MASK": 4}),
        ("dual", {"IPCLB200_FP_MASK": 6}),
        ("dual2", {"IPCLB200_DUAL2": "2,1"}),
        ("dual2", {"IPCLB200_DUAL2": "1,2"}),
        ("dual2", {"IPCLB200_DUAL2": "2,0"}),
        ("dual2", {"IPCLB200_DUAL2": "1,0"}),
        ("dual2", {"IPCLB200_DUAL2": "0,1"}),
        ("dual2", {"IPCLB200_DUAL2": "0,2"}),
    ]
    if args.configs:
        keep = set(int(x) for x in args.configs.split(","))
        configs = [c for i, c in enumerate(configs) if i in keep]
    print("residues vs pow() on %d ciphertexts" % args.small, flush=True)
    want_res = [[pow(c, p - 1, p * p), pow(c, q - 1, q * q)] for c in cts]
    good = []
    for mode, env in configs:
        set_mode(mode, **env)
        try:
            x = sk.crt_residues(ct_small)
        except Exception as e:  # noqa: BLE001
            print("  %-6s %-28s ERROR %r" % (mode, env, e), flush=True)
            ok_all = False
            continue
        bad = [(i, s) for i in range(args.small) for s in (0, 1)
               if int.from_bytes(x[i, s].tobytes(), "little") != want_res[i][s]]
        print("  %-6s %-28s %s" % (mode, env, "ok" if not bad else "MISMATCH %s" % bad[:20]),
              flush=True)
        if bad:
            ok_all = False
        else:
            good.append((mode, env))

    # timing, device resident
    B = args.count
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    pt = random_limbs(rng, B, 64, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, B, 32)
    d_pt = torch.from_numpy(pt.view(np.int32)).to(dev)
    d_r = torch.from_numpy(r.view(np.int32)).to(dev)
    d_ct = torch.empty((B, 128), dtype=torch.int32, device=dev)
    d_dt = torch.empty((B, 64), dtype=torch.int32, device=dev)
    pk.encrypt_dev(d_pt.data_ptr(), 64, d_r.data_ptr(), 32, B, d_ct.data_ptr(), stream)
    torch.cuda.synchronize()
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    print("decrypt timing, batch %d (ms per launch, device resident)" % B, flush=True)
    for mode, env in good:
        set_mode(mode, **env)
        ms = []
        clk = []
        for rep in range(args.reps + 1):
            d_dt.zero_()
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), stream)
            e1.record()
            time.sleep(0.03)
            clk.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            e1.synchronize()
            if rep:
                ms.append(e0.elapsed_time(e1))
        same = bool(torch.equal(d_dt, d_pt))
        print("  %-6s %-28s %s  best %.2f ms  (%s)  sm %s MHz  %.0f W  roundtrip %s" % (
            mode, env, "", min(ms), ", ".join("%.2f" % m for m in ms), clk[-1], pw,
            "ok" if same else "WRONG"), flush=True)
        if not same:
            ok_all = False
    names = {0: "IMAD.WIDE on all 32 warps/SM", 1: "DFMA on all 32 warps/SM",
             2: "16 warps IMAD.WIDE + 16 warps DFMA (both kinds on every sub-partition)",
             3: "the 16 IMAD.WIDE warps of mode 2 alone",
             4: "the 16 DFMA warps of mode 2 alone"}
    print("pipe overlap probe (ms, best of 3)")
    for mode in range(5):
        print("  mode %d  %8.3f ms   %s" % (mode, capi.pipe_mix(mode), names[mode]), flush=True)
    print("PROBE", "OK" if ok_all else "FAILED")


if __name__ == "__main__":
    main()
