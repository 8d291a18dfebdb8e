"""worker for tests/test_gpu_multi.py: N-rank NCCL run of an observation-sharded fit"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from gslnls_b200 import Model, Problem  # noqa: E402
from gslnls_b200.distributed import init_comm_from_torch, shard_bounds  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(os.environ.get("GSLNLS_TEST_N", "2000003"))
    lo, hi = shard_bounds(n, rank, world)
    x, y = bench.synth_rows(lo, hi, n)
    m = Model(bench.FORMULA_RHS, ["A", "lam", "b"], ["x"], jac=True, fvv=True)
    pb = Problem(m, hi - lo, False, local).upload([x], y)
    comm = init_comm_from_torch(local)
    pb.set_comm(comm)
    theta = [4.0, 1.3, 0.9]
    pk = pb.eval_packet(theta)
    out = {"rank": rank, "packet": pk.tolist(), "fits": {}, "p2p": int(comm.has_peer_memory)}
    for alg in ("lm", "lmaccel", "dogleg"):
        f = pb.fit(list(bench.START), algorithm=alg)
        out["fits"][alg] = {"par": f["par"].tolist(), "ssr": f["ssr"], "niter": f["niter"], "conv": f["conv"],
                            "n": int(f["n"])}
    with open(os.path.join(os.environ["GSLNLS_TEST_OUT"], "rank%d.json" % rank), "w") as fh:
        json.dump(out, fh)
    pb.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
