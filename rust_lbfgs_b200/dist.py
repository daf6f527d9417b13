"""Multi-GPU plumbing: one process per GPU, contiguous 1-D shards, one NCCL rank per process.

torch.distributed is used only to ship the NCCL unique id from rank 0 and to reduce timings; the
data path is the solver's own scalar all-reduce inside liblbfgsb200.so.  Every
function here except `Comm` works on CPU with the gloo backend.
"""
import ctypes as C

from . import _lib


def shard_range(n_global, rank, world_size, granule=2):
    """Contiguous [lo, hi) of rank's shard.  Boundaries are multiples of `granule` (2 keeps the
    Rosenbrock pairs (2i, 2i+1) of src/lib.rs:84 on one rank); the last rank takes the remainder."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    units = n_global // granule
    base, extra = divmod(units, world_size)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    lo, hi = lo_u * granule, hi_u * granule
    if rank == world_size - 1:
        hi = n_global
    return lo, hi


def owl_range_local(start, end, n_global, lo, hi):
    """Intersection of the global OWL-QN range [start, end) (end None => n_global, clamped,
    src/orthantwise.rs:59-67) with this rank's [lo, hi), in local indices; None if empty."""
    e = n_global if end is None else min(end, n_global)
    a, b = max(start, lo), min(e, hi)
    return None if a >= b else (a - lo, b - lo)


def broadcast_unique_id(make_id, group=None):
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if rank == 0:
        raw = make_id()
        assert len(raw) == _lib.UNIQUE_ID_BYTES
        t = torch.tensor(list(raw), dtype=torch.uint8)
    else:
        t = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


def max_over_ranks(value, group=None):
    """Max of a python float over ranks (device-timed milliseconds)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, group=None):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def _new_unique_id():
    buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
    st = _lib.lib().lbfgsb200_comm_unique_id(buf)
    if st != 0:
        raise RuntimeError("lbfgsb200_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


class Comm:
    """The solver's NCCL communicator for this process's GPU (needs a GPU)."""

    def __init__(self, rank, world_size, device, unique_id=None, group=None):
        self.rank, self.world_size, self.device = rank, world_size, device
        if unique_id is None:
            unique_id = broadcast_unique_id(_new_unique_id, group)
        out = C.c_void_p()
        st = _lib.lib().lbfgsb200_comm_create(unique_id, rank, world_size, device, C.byref(out))
        if st != 0:
            raise RuntimeError(f"lbfgsb200_comm_create failed: {_lib.STATUS_NAMES.get(st, st)}")
        self._handle = out

    @property
    def transport(self):
        """"peer_mailboxes" (exchange fused into the kernels over NVLink peer stores) or "nccl_allreduce"."""
        return "peer_mailboxes" if _lib.lib().lbfgsb200_comm_transport(self._handle) == 1 else "nccl_allreduce"

    def allreduce_sum_(self, tensor):
        import torch
        st = _lib.lib().lbfgsb200_comm_allreduce_sum(self._handle, tensor.data_ptr(), tensor.numel(),
                                                     int(torch.cuda.current_stream().cuda_stream))
        if st != 0:
            raise RuntimeError("allreduce failed")
        return tensor

    def close(self):
        if self._handle:
            _lib.lib().lbfgsb200_comm_destroy(self._handle)
            self._handle = None
