#!/usr/bin/env python3
"""Fixed-base table width of the DJN encrypt (IPCLB200_COMB_WINDOW, read when the
table is built): build time of the wide table, device-resident encrypt time of
one batch, bit-exact check of a sample against the oracle.  One JSON line.
`python tools/comb_probe.py <key bits> <count>`; run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402


def main():
    bits = sys.argv[1] if len(sys.argv) > 1 else "2048"
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[bits].items()}
    p, q = sorted((k["p"], k["q"]))
    NL = int(bits) // 32
    capi.init(0)
    os.environ["IPCLB200_COMB_SYNC"] = "1"
    rng = np.random.default_rng(9)
    nl, hs = to_limbs(p * q, NL), to_limbs(k["hs"], 2 * NL)
    pk = capi.PubKey(nl, hs, int(bits) // 2)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, NL // 2)
    d_pt = torch.from_numpy(pt.view(np.int32)).cuda()
    d_r = torch.from_numpy(r.view(np.int32)).cuda()
    d_ct = torch.zeros((count, 2 * NL), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()

    def enc():
        pk.encrypt_dev(d_pt.data_ptr(), NL, d_r.data_ptr(), NL // 2, count, d_ct.data_ptr(),
                       s.cuda_stream)

    free0 = torch.cuda.mem_get_info()[0]
    t0 = time.perf_counter()
    enc()          # builds the starter table, then (count >= 8192) the wide one
    s.synchronize()
    first_ms = (time.perf_counter() - t0) * 1e3
    enc()
    s.synchronize()
    table_mb = (free0 - torch.cuda.mem_get_info()[0]) / 2 ** 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        enc()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    import oracle as orc
    S = 512
    ok = bool(np.array_equal(d_ct[:S].cpu().numpy().view(np.uint32),
                             orc.encrypt(nl, hs, pt[:S], r[:S])))
    print(json.dumps({"bits": bits, "count": count,
                      "IPCLB200_COMB_WINDOW": os.environ.get("IPCLB200_COMB_WINDOW", "default"),
                      "IPCLB200_COMB_MAX_MB": os.environ.get("IPCLB200_COMB_MAX_MB", "default"),
                      "first_call_ms": round(first_ms, 1), "hbm_in_use_mb": round(table_mb),
                      "encrypt_ms": round(ms, 3), "enc_per_s": round(count / ms * 1e3),
                      "oracle_ok": ok}), flush=True)


if __name__ == "__main__":
    main()
