"""Latency of small batches with the default layout and the wide (4 limbs per
lane) layout, device-resident, CUDA events, best of 5.  Results of the two
layouts are compared bit for bit."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402

capi.init(0)
with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
    K = {b: {k: int(v, 16) for k, v in d.items()} for b, d in json.load(f).items()}
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(4)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


counts = [256, 1024, 2048, 3072, 4096, 6144, 8192, 12288, 16384]
for bits in ("2048", "1024"):
    k = K[bits]
    p, q = sorted((k["p"], k["q"]))
    NL = int(bits) // 32
    pk = capi.PubKey(to_limbs(p * q, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
    sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
    nsq = to_limbs((p * q) ** 2, 2 * NL)
    for B in counts:
        d_pt = t(random_limbs(rng, B, NL, top_mask=0x3FFFFFFF))
        d_r = t(random_limbs(rng, B, NL // 2))
        d_e = t(random_limbs(rng, B, NL))
        d_ct = torch.empty((B, 2 * NL), dtype=torch.int32, device=dev)
        d_dt = torch.empty((B, NL), dtype=torch.int32, device=dev)
        d_o = torch.empty((B, 2 * NL), dtype=torch.int32, device=dev)
        res = {}
        for wide in ("0", "2", "1"):
            os.environ["IPCLB200_WIDE"] = wide
            enc = timed(lambda: pk.encrypt_dev(d_pt.data_ptr(), NL, d_r.data_ptr(), NL // 2, B,
                                               d_ct.data_ptr(), st))
            ct = d_ct.clone()
            dec = timed(lambda: sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), st))
            assert torch.equal(d_dt, d_pt)
            mul = timed(lambda: capi.modexp_dev(d_ct.data_ptr(), d_e.data_ptr(), nsq, NL, 32 * NL, B,
                                                d_o.data_ptr(), st), reps=2)
            res[wide] = (enc, dec, mul, ct, d_o.clone())
        for w in ("1", "2"):
            assert torch.equal(res["0"][3], res[w][3]) and torch.equal(res["0"][4], res[w][4])
        print("key %s batch %5d | encrypt %7.3f / %7.3f / %7.3f ms | decrypt %7.3f / %7.3f / %7.3f ms | ct*pt %8.3f / %8.3f / %8.3f ms   (default / mid / wide)"
              % (bits, B, res["0"][0], res["2"][0], res["1"][0], res["0"][1], res["2"][1], res["1"][1],
                 res["0"][2], res["2"][2], res["1"][2]), flush=True)
