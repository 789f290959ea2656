"""bench.py contract checks that need no GPU: the reference arm prints exactly
one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl",
                        "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "16"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl",
                        "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_mirrors_of_library_choices():
    """bench.py restates two host-side choices of the library to count the
    multiplies a launch executes (roofline.executed_frac): the sliding-window
    width of the two-digit decrypt (ipclb200_privkey_create) and the width of
    the wide fixed-base table (comb_pick_window).  Pin them on the golden keys."""
    sys.path.insert(0, ROOT)
    import bench
    for bits, hw, cw in ((1024, 5, 18), (2048, 6, 17), (3072, 6, 15)):
        p, q, _ = bench.load_key(bits)
        assert bench.hensel_window(p, q) == hw
        assert bench.comb_window(bits) == cw
    p, q, _ = bench.load_key(2048)
    nsq, nmul = bench.sliding_counts(p - 1, 6)
    assert (nsq, nmul) == (1019, 142)
    assert bench.hensel_executed_macs(p, q) == 10230496
    # generic algorithm of SURVEY.md section 8(d)
    assert bench.MAC_DECRYPT == 20854656 and bench.MAC_ENCRYPT == 1263 * (2 * 128 * 128 + 128)
