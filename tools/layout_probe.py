"""Timing probe of the lane layouts of the two-digit CRT decrypt on one GPU:
device-resident, CUDA events on the launching stream, one JSON line per (batch
size, layout).  `python tools/layout_probe.py <key bits> <count> [<count> ...]`,
layouts from PROBE_LAYOUTS (default "0,1,-2": 16 limbs x 2 lanes, 8 x 4, one task
per thread).  profiles/r02_layout_sweep.jsonl and r02_layout_other_keys.jsonl are
its output; the cost model pick_hensel_spread (csrc/host_common.hpp) is fitted to
them.  Run under gpurun."""
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
    counts = [int(a) for a in sys.argv[2:]] or [65536]
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a: int(b, 16) for a, b in json.load(f)[bits].items()}
    p, q = sorted((k["p"], k["q"]))
    n = p * q
    NL = int(bits) // 32
    capi.init(0)
    rng = np.random.default_rng(5)
    pk = capi.PubKey(to_limbs(n, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    stream = torch.cuda.Stream()
    for count in counts:
        pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
        r = random_limbs(rng, count, NL // 2)
        ct = pk.encrypt(pt, r)
        d_ct = torch.from_numpy(ct.view(np.int32)).cuda()
        d_pt = torch.zeros((count, NL), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        configs = [{"IPCLB200_HENSEL_SPREAD": v} for v in
                   os.environ.get("PROBE_LAYOUTS", "0,1,-2").split(",")]
        for env in configs:
            for a in ("IPCLB200_HENSEL_ROWS", "IPCLB200_HENSEL_BLOCKS", "IPCLB200_HENSEL_SPREAD",
                      "IPCLB200_HENSEL_W64"):
                os.environ.pop(a, None)
            os.environ.update(env)
            with torch.cuda.stream(stream):
                d_pt.zero_()
                for _ in range(2):
                    sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
                stream.synchronize()
                ok = bool(np.array_equal(d_pt.cpu().numpy().view(np.uint32), pt))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                reps = 3
                for _ in range(reps):
                    sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), stream.cuda_stream)
                b.record(stream)
                stream.synchronize()
            ms = a.elapsed_time(b) / reps
            if "IPCLB200_HENSEL_WINDOW" in os.environ:   # read at key creation
                env = dict(env, IPCLB200_HENSEL_WINDOW=os.environ["IPCLB200_HENSEL_WINDOW"])
            print(json.dumps({"bits": bits, "count": count, "env": env, "ms": round(ms, 3),
                              "dec_per_s": round(count / ms * 1e3), "ok": ok}), flush=True)


if __name__ == "__main__":
    main()
