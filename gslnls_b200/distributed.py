"""Observation sharding across the GPUs of one box (one process per GPU, torchrun-style launch).

The path shards by observation: J^T J, J^T f and f^T f are sums over independent rows, so rank r
keeps rows [lo_r, hi_r) resident in its HBM for the whole fit and the only exchange is one
all-reduce of the p(p+1)/2 + p + 1 double packet per pass.  torch.distributed is used
for rendezvous only (shipping the 128-byte communicator id); the packet exchange itself happens
inside libgslnls_b200.so on the solver's own stream.
"""
import ctypes as C

from . import _lib


def shard_bounds(n, rank, world):
    """contiguous, 2-aligned row ranges (16-byte alignment keeps the vector loads on every shard)"""
    per = (n + world - 1) // world
    per += per & 1
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


class Comm:
    def __init__(self, rank=0, world=1, device=0, id_bytes=None):
        h = C.c_void_p()
        buf = C.create_string_buffer(id_bytes, 128) if id_bytes is not None else None
        _lib.check(_lib.lib().gslnls_comm_create(buf, rank, world, device, C.byref(h)))
        self.handle, self.rank, self.world = h, rank, world

    @property
    def has_peer_memory(self):
        """True when packets travel through NVLink peer-memory mailboxes (no collective call per pass)"""
        return bool(_lib.lib().gslnls_comm_has_peer_memory(self.handle))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().gslnls_comm_get_unique_id(buf))
        return buf.raw

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().gslnls_comm_free(self.handle)
            self.handle = None


def exchange_unique_id(make_id, dist=None):
    """rank 0 creates the id, every rank returns the same 128 bytes (broadcast over torch.distributed;
    works with the gloo backend on CPU and with nccl on GPUs)"""
    import torch
    import torch.distributed as td
    dist = dist or td
    rank = dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.tensor(list(make_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def init_comm_from_torch(device):
    """build a Comm for the current torch.distributed process group (no-op group of 1 otherwise)"""
    import torch.distributed as td
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return Comm(0, 1, device)
    idb = exchange_unique_id(Comm.unique_id)
    return Comm(td.get_rank(), td.get_world_size(), device, idb)
