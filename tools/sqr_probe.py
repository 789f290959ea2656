"""GPU probe for MontSqr::sqr: single squarings against exact integers, then
decrypt residues with IPCLB200_DECRYPT=sqr against pow(), then timing."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from pailliercryptolib_b200 import capi  # noqa: E402
from pailliercryptolib_b200.limbs import (batch_from_limbs, batch_to_limbs,  # noqa: E402
                                          random_limbs, to_limbs)

capi.init(0)
rng = np.random.default_rng(64)
R = 1 << 2048
ok = True
LAYOUT = sys.argv[1] if len(sys.argv) > 1 else "1"
MODE = "sqr2" if LAYOUT == "2" else "sqr"
os.environ["IPCLB200_DEBUG_SQR_LAYOUT"] = LAYOUT
for trial in range(3):
    mod = random_limbs(rng, 1, 64)
    mod[0, 0] |= 1
    mod[0, -1] |= 0x80000000
    n = batch_from_limbs(mod)[0]
    vals = [0, 1, R - 1, n - 1, int("ffffffff00000000" * 32, 16)] + \
        [int.from_bytes(rng.bytes(256), "little") for _ in range(95)]
    a = batch_to_limbs(vals, 64)
    s, m = capi.debug_montsqr(a, mod[0])
    S, Mv = batch_from_limbs(s), batch_from_limbs(m)
    bad_mul = [i for i, (x, v) in enumerate(zip(Mv, vals)) if (x * R - v * v) % n or x >= R]
    bad = [i for i, (x, v) in enumerate(zip(S, vals)) if (x * R - v * v) % n or x >= R]
    print("trial %d: mul wrong %s, sqr wrong %d of %d %s" % (trial, bad_mul[:5], len(bad), len(vals), bad[:10]),
          flush=True)
    if bad:
        ok = False
        i = bad[0]
        # localise: compare with the exact product pieces
        v = vals[i]
        sq = v * v
        W, H = sq % R, sq // R
        npr = (-pow(n, -1, R)) % R
        V = (W + ((W * npr) % R) * n) // R
        want = V + H
        if want >= R:
            want -= n
        print("  a    = %x" % v)
        print("  got  = %x" % S[i])
        print("  want = %x" % want)
        d = S[i] ^ want
        print("  differing bits: lowest %d highest %d" % ((d & -d).bit_length() - 1, d.bit_length() - 1))
        print("  got - want = %x" % (S[i] - want))
        break

if ok:
    with open(os.path.join(ROOT, "tests", "golden", "keys.json")) as f:
        k = {a_: int(b, 16) for a_, b in json.load(f)["2048"].items()}
    p, q = sorted((k["p"], k["q"]))
    nsq = (p * q) ** 2
    sk = capi.PrivKey(to_limbs(p, 32), to_limbs(q, 32))
    pk = capi.PubKey(to_limbs(p * q, 64), to_limbs(k["hs"], 128), 1024)
    count = 3001
    cts = [int.from_bytes(rng.bytes(512), "little") % nsq for _ in range(count)]
    cts[:5] = [0, 1, nsq - 1, p * 12345, q * q * 3 % nsq]
    ct = batch_to_limbs(cts, 128)
    os.environ["IPCLB200_DECRYPT"] = MODE
    os.environ["IPCLB200_WIDE"] = "0"
    x = sk.crt_residues(ct)
    bad = [(i, s_) for i in range(300) for s_ in (0, 1)
           if int.from_bytes(x[i, s_].tobytes(), "little") != pow(cts[i], (p, q)[s_] - 1, (p, q)[s_] ** 2)]
    os.environ["IPCLB200_DECRYPT"] = "int"
    x0 = sk.crt_residues(ct)
    for extra in ("k32s", "k32s2"):
        os.environ["IPCLB200_DECRYPT"] = extra
        same = bool(np.array_equal(sk.crt_residues(ct), x0))
        print("residues %s equal to the int kernel: %s" % (extra, same), flush=True)
        ok = ok and same
    print("residues: %d wrong of 600 vs pow(); equal to the int kernel on all %d: %s" % (
        len(bad), count, bool(np.array_equal(x, x0))), flush=True)
    ok = ok and not bad and np.array_equal(x, x0)

if ok:
    import torch
    B = 65536
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    pt = random_limbs(rng, B, 64, top_mask=0x3FFFFFFF)
    d_pt = torch.from_numpy(pt.view(np.int32)).to(dev)
    d_r = torch.from_numpy(random_limbs(rng, B, 32).view(np.int32)).to(dev)
    d_ct = torch.empty((B, 128), dtype=torch.int32, device=dev)
    d_dt = torch.empty((B, 64), dtype=torch.int32, device=dev)
    pk.encrypt_dev(d_pt.data_ptr(), 64, d_r.data_ptr(), 32, B, d_ct.data_ptr(), st)
    for mode in ("int", MODE, "k32", "k32s", "k32s2"):
        os.environ["IPCLB200_DECRYPT"] = mode
        ms = []
        for rep in range(4):
            d_dt.zero_()
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            sk.decrypt_dev(d_ct.data_ptr(), B, d_dt.data_ptr(), st)
            e1.record()
            e1.synchronize()
            if rep:
                ms.append(e0.elapsed_time(e1))
        print("decrypt %s: %s ms, round trip %s" % (mode, ["%.2f" % v for v in ms],
                                                    bool(torch.equal(d_dt, d_pt))), flush=True)
print("SQR PROBE", "OK" if ok else "FAILED")
