"""bench.py contract checks that need no GPU: the reference arm runs the oracle port on the host cores and
prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5",
                          "--warmup", "1", "--cpu-sample", "100000"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] >= 5 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
    assert "configs[2]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_rows_are_shard_independent():
    """any shard regenerates exactly the doubles of the full data set (counter-based stream per chunk)"""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    from gslnls_b200.distributed import shard_bounds
    n = 3 * bench.CHUNK + 12345
    x, y = bench.synth_rows(0, n, n)
    for world in (2, 3, 8):
        xs, ys = [], []
        for r in range(world):
            lo, hi = shard_bounds(n, r, world)
            assert lo % 2 == 0
            a, b = bench.synth_rows(lo, hi, n)
            xs.append(a)
            ys.append(b)
        assert np.array_equal(np.concatenate(xs), x) and np.array_equal(np.concatenate(ys), y)


def test_reference_arm_of_the_sparse_configurations():
    """--config sparse / penalty500: the CPU arm is oracle/sparse.py, one JSON line each"""
    for cfg, extra, unit in (("sparse", ["--cpu-sample", "40000"], "iterations/s"), ("penalty500", ["--n", "60"], "fits/s")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg,
                              "--steps", "2", "--warmup", "1"] + extra, capture_output=True, text=True, timeout=600,
                             cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, lines
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["gpu_launches"] == 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
        assert d["e2e"]["value"] == d["value"] and d["config"]["final"]["ssr"] > 0
