"""Host-side multi-GPU logic on CPU: shard arithmetic, and scatter -> per-rank
work -> gather -> max-over-ranks under torch.distributed with the gloo backend
at world size 2 (the NCCL path runs the same code on GPUs)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_exactly():
    from pailliercryptolib_b200.sharding import shard_range, shard_sizes
    for count in [0, 1, 7, 8, 9, 65536, 262144, 100003]:
        for world in [1, 2, 3, 4, 8]:
            spans = [shard_range(count, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = shard_sizes(count, world)
            assert sum(sizes) == count and max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from pailliercryptolib_b200 import sharding
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import oracle as orc
from pailliercryptolib_b200.limbs import random_limbs, to_limbs

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
L, count = 16, 37                     # ragged: 18 + 19
rng = np.random.default_rng(3)
mod = random_limbs(rng, 1, L); mod[0, 0] |= 1
base = random_limbs(rng, count, L); exp = random_limbs(rng, count, 2)
as_t = lambda a: torch.from_numpy(a.view(np.int32).copy())
full_b = as_t(base) if rank == 0 else None
full_e = as_t(exp) if rank == 0 else None
lb = sharding.scatter_rows(full_b, count, L, torch.int32, "cpu")
le = sharding.scatter_rows(full_e, count, 2, torch.int32, "cpu")
s, e = sharding.shard_range(count, world, rank)
assert lb.shape[0] == e - s
# each rank works on its own shard (the oracle stands in for the kernel here)
out = orc.modexp(lb.numpy().view(np.uint32), le.numpy().view(np.uint32), mod, shared_mod=True)
res = sharding.gather_rows(as_t(out), count)
t = sharding.max_over_ranks([float(rank + 1), 5.0 - rank], "cpu")
assert t == [float(world), 5.0]
if rank == 0:
    want = orc.modexp(base, exp, mod, shared_mod=True)
    assert np.array_equal(res.numpy().view(np.uint32), want)
    print("SHARDING_OK")
dist.destroy_process_group()
"""


def test_scatter_work_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SHARDING_OK" in r.stdout
