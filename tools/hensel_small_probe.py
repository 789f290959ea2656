"""Two-digit decrypt at small and medium batches: lane-spread layouts
(IPCLB200_HENSEL_SPREAD = 0/1/2: one task over 2/4/8 lanes at a 2048-bit key)
against the full-width kernel.  Device-resident, CUDA events.  Run under gpurun."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402


def main():
    bits = sys.argv[1] if len(sys.argv) > 1 else "2048"
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[bits].items()}
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    capi.init(0)
    rng = np.random.default_rng(5)
    top = 16384
    pt = random_limbs(rng, top, NL, top_mask=0x3FFFFFFF)
    r = random_limbs(rng, top, NL // 2)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    ct = pk.encrypt(pt, r)
    stream = torch.cuda.Stream()
    d_ct = torch.from_numpy(ct.view(np.int32)).cuda()
    d_pt = torch.zeros((top, NL), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for count in (16, 256, 1024, 2048, 4096, 8192, 16384):
        for mode, spread in (("int", None), ("hensel", "0"), ("hensel", "1"), ("hensel", "2"),
                             ("hensel", "auto")):
            os.environ["IPCLB200_DECRYPT"] = mode
            os.environ.pop("IPCLB200_HENSEL_SPREAD", None)
            if spread not in (None, "auto"):
                os.environ["IPCLB200_HENSEL_SPREAD"] = spread
            with torch.cuda.stream(stream):
                for _ in range(2):
                    sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
                stream.synchronize()
                ok = bool(np.array_equal(d_pt[:count].cpu().numpy().view(np.uint32), pt[:count]))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                reps = 5
                for _ in range(reps):
                    sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
                b.record(stream)
                stream.synchronize()
            ms = a.elapsed_time(b) / reps
            print(json.dumps({"bits": bits, "count": count, "mode": mode, "spread": spread,
                              "ms": round(ms, 3), "dec_per_s": round(count / ms * 1e3),
                              "ok": ok}), flush=True)


if __name__ == "__main__":
    main()
