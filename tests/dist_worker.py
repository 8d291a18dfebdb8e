"""worker for tests/test_distributed_cpu.py: world_size-2 gloo run of the host-side sharding logic"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from gslnls_b200.distributed import exchange_unique_id, shard_bounds  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 200_003
    lo, hi = shard_bounds(n, rank, world)
    x, y = bench.synth_rows(lo, hi, n)
    theta = [4.0, 1.3, 0.9]
    pk = O.eval_packet("exp3", y, theta, x=x)
    t = torch.tensor(pk)
    dist.all_reduce(t)  # the one collective of the path: sum of the per-shard packets
    idb = exchange_unique_id(lambda: bytes(range(128)))
    bounds = [None] * world
    dist.all_gather_object(bounds, (lo, hi))
    out = {"rank": rank, "id_ok": idb == bytes(range(128)), "bounds": bounds, "packet": t.tolist()}
    if rank == 0:
        xf, yf = bench.synth_rows(0, n, n)
        full = O.eval_packet("exp3", yf, theta, x=xf, longdouble=True)
        out["full"] = full.tolist()
        out["shard_rows_match"] = bool(np.array_equal(xf[lo:hi], x) and np.array_equal(yf[lo:hi], y))
    with open(os.path.join(os.environ["GSLNLS_TEST_OUT"], "rank%d.json" % rank), "w") as fh:
        json.dump(out, fh)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
