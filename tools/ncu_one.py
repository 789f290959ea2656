"""One CRT-decrypt configuration, a few launches: the target of `ncu -k
regex:decrypt_hensel` (layout chosen by the IPCLB200_HENSEL_* environment)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402

bits = sys.argv[1] if len(sys.argv) > 1 else "2048"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
    k = {a: int(b, 16) for a, b in json.load(f)[bits].items()}
p, q = sorted((k["p"], k["q"]))
NL = int(bits) // 32
capi.init(0)
rng = np.random.default_rng(5)
pt = random_limbs(rng, count, NL, top_mask=0x3FFFFFFF)
r = random_limbs(rng, count, NL // 2)
pk = capi.PubKey(to_limbs(p * q, NL), to_limbs(k["hs"], 2 * NL), int(bits) // 2)
sk = capi.PrivKey(to_limbs(p, NL // 2), to_limbs(q, NL // 2))
ct = pk.encrypt(pt, r)
d_ct = torch.from_numpy(ct.view(np.int32)).cuda()
d_pt = torch.zeros((count, NL), dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
for _ in range(4):
    sk.decrypt_dev(d_ct.data_ptr(), count, d_pt.data_ptr(), s.cuda_stream)
s.synchronize()
print("ok", bool(np.array_equal(d_pt.cpu().numpy().view(np.uint32), pt)))
