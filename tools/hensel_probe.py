"""Timing probe of the CRT-decrypt kernels on one GPU: two-digit (Hensel) kernel
at the row-unroll / blocks-per-SM variants against the full-width kernel.
Device-resident, CUDA events on the library stream; prints one JSON line per
configuration.  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402


def main():
    bits = sys.argv[1] if len(sys.argv) > 1 else "2048"
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[bits].items()}
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    capi.init(0)
    rng = np.random.default_rng(5)
    pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, count, NL // 2)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    ct = pk.encrypt(pt, r)
    stream = torch.cuda.Stream()
    d_ct = torch.from_numpy(ct.view(np.int32)).cuda()
    d_pt = torch.zeros((count, NL), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    configs = [("int", {}), ("hensel", {"IPCLB200_HENSEL_ROWS": "8"}),
               ("hensel", {"IPCLB200_HENSEL_ROWS": "4"}),
               ("hensel", {"IPCLB200_HENSEL_ROWS": "16"}),
               ("hensel", {"IPCLB200_HENSEL_ROWS": "8", "IPCLB200_HENSEL_BLOCKS": "2"}),
               ("hensel", {"IPCLB200_HENSEL_ROWS": "8", "IPCLB200_HENSEL_BLOCKS": "1"})]
    for mode, env in configs:
        os.environ["IPCLB200_DECRYPT"] = mode
        for a in ("IPCLB200_HENSEL_ROWS", "IPCLB200_HENSEL_BLOCKS"):
            os.environ.pop(a, None)
        os.environ.update(env)
        with torch.cuda.stream(stream):
            for _ in range(2):
                sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            ok = bool(np.array_equal(d_pt.cpu().numpy().view(np.uint32), pt))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            reps = 5
            for _ in range(reps):
                sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
            b.record(stream)
            stream.synchronize()
        ms = a.elapsed_time(b) / reps
        print(json.dumps({"bits": bits, "count": count, "mode": mode, "env": env,
                          "ms": round(ms, 3), "dec_per_s": round(count / ms * 1e3),
                          "ok": ok}), flush=True)


if __name__ == "__main__":
    main()
