"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo backend, 127.0.0.1 rendezvous.
Checks the shard map, that shards regenerate identical rows, that the path's single collective (sum of
per-shard packets) reproduces the full packet, and the 128-byte id broadcast used for NCCL set-up."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from gslnls_b200.distributed import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    for n in (1, 2, 3, 25, 1000, 100_000_000, 100_000_001):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert lo == prev and lo <= hi <= n
                assert lo % 2 == 0 or lo == n  # 16-byte aligned shard starts keep the vector loads
                prev = hi
            assert prev == n


def test_two_rank_gloo_packet_sum(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, GSLNLS_TEST_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(ROOT, "tests", "dist_worker.py")], env=env, timeout=300)
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["id_ok"] and r1["id_ok"]
    assert r0["shard_rows_match"]
    assert r0["bounds"] == r1["bounds"] and r0["bounds"][0][1] == r0["bounds"][1][0]
    assert r0["packet"] == r1["packet"]  # every rank holds bitwise the same reduced packet
    full, got = np.array(r0["full"]), np.array(r0["packet"])
    assert np.max(np.abs(got - full) / np.abs(full)) < 1e-12
