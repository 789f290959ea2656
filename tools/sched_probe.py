"""Timing of the shared-exponent paths with and without the sliding-window
schedule (IPCLB200_NO_SCHED=1): non-DJN encrypt (r^n), RAW decrypt (ct^lambda),
ct * scalar.  2048-bit key, device-resident, CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
capi.init(0)
with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
    k = {a: int(b, 16) for a, b in json.load(f)["2048"].items()}
p, q = sorted((k["p"], k["q"]))
n = p * q
pk = capi.PubKey(to_limbs(n, 64))
sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
rng = np.random.default_rng(3)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
pt = random_limbs(rng, B, 64, top_mask=0x3FFFFFFF)
r = random_limbs(rng, B, 64, top_mask=0x3FFFFFFF)
r[:, 0] |= 1
d_pt, d_r = t(pt), t(r)
d_ct = torch.empty((B, 128), dtype=torch.int32, device=dev)
d_dt = torch.empty((B, 64), dtype=torch.int32, device=dev)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for label, env in (("schedule", None), ("fixed window", "1")):
    if env:
        os.environ["IPCLB200_NO_SCHED"] = env
    enc = timed(lambda: pk.encrypt_dev(d_pt.data_ptr(), 64, d_r.data_ptr(), 64, B, d_ct.data_ptr(), st))
    raw = timed(lambda: sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), st, use_crt=False))
    ok = bool(torch.equal(d_dt, d_pt))
    print("%-13s non-DJN encrypt %.1f ms (%.0f/s)   RAW decrypt %.1f ms (%.0f/s)   round trip %s" % (
        label, enc, B / enc * 1e3, raw, B / raw * 1e3, "ok" if ok else "WRONG"), flush=True)
