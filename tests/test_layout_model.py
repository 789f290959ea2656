"""The cost model that picks the lane layout of the CRT decrypt
(pick_hensel_spread, csrc/host_common.hpp, exported as the host-only diagnostic
ipclb200_decrypt_layout) against the measured sweep it was fitted to
(profiles/r02_layout_sweep.jsonl: 2048-bit key, one B200, layouts 0 / 1 / -2 at 21
batch sizes): the pick must be within 5 % of the fastest measured layout at every
size, and the thread-per-task layout must be the pick at the headline batch."""
import collections
import json
import os

from pailliercryptolib_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sweep():
    d = collections.defaultdict(dict)
    with open(os.path.join(ROOT, "profiles", "r02_layout_sweep.jsonl")) as f:
        for line in f:
            if line.startswith("{"):
                r = json.loads(line)
                assert r["ok"]
                d[r["count"]][int(r["env"]["IPCLB200_HENSEL_SPREAD"])] = r["ms"]
    return d


def test_pick_is_within_5_percent_of_the_measured_best():
    sweep = _sweep()
    assert len(sweep) >= 20
    for count, times in sorted(sweep.items()):
        pick = capi.decrypt_layout(count, 32, 148)
        if pick == 2:  # not in the sweep (only competitive below 4096)
            continue
        assert pick in times, (count, pick)
        assert times[pick] <= 1.05 * min(times.values()), (count, pick, times)


def test_headline_batch_runs_one_task_per_thread():
    assert capi.decrypt_layout(65536, 32, 148) == -2
    assert capi.decrypt_layout(262144, 32, 148) == -2
    # small batches spread a task over more lanes; other key sizes never pick -2
    assert capi.decrypt_layout(16, 32, 148) == 2
    assert capi.decrypt_layout(8192, 32, 148) == 0
    for pl in (16, 48, 64):
        for count in (8, 1000, 65536):
            assert capi.decrypt_layout(count, pl, 148) in (0, 1)
    # 3072-bit keys: 24 x 2 lanes (8 warps/SM) only while the launch is one round
    assert capi.decrypt_layout(8192, 48, 148) == 0
    assert capi.decrypt_layout(32768, 48, 148) == 1
    assert capi.decrypt_layout(262144, 48, 148) == 1
