"""one decrypt launch with the integer kernel and one with the symmetric-
squaring kernel, for ncu (batch 16384, 2048-bit key)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import random_limbs, to_limbs  # noqa: E402

B = 16384
capi.init(0)
with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
    k = {a: int(b, 16) for a, b in json.load(f)["2048"].items()}
p, q = sorted((k["p"], k["q"]))
pk = capi.PubKey(to_limbs(p * q, 64), to_limbs(k["hs"], 128), 1024)
sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
rng = np.random.default_rng(1)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d_pt = torch.from_numpy(random_limbs(rng, B, 64, top_mask=0x3FFFFFFF).view(np.int32)).to(dev)
d_r = torch.from_numpy(random_limbs(rng, B, 32).view(np.int32)).to(dev)
d_ct = torch.empty((B, 128), dtype=torch.int32, device=dev)
d_dt = torch.empty((B, 64), dtype=torch.int32, device=dev)
pk.encrypt_dev(d_pt.data_ptr(), 64, d_r.data_ptr(), 32, B, d_ct.data_ptr(), st)
for mode in sys.argv[1:] or ["int", "sqr"]:
    os.environ["IPCLB200_DECRYPT"] = mode
    sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(d_dt, d_pt), mode
print("ok")
