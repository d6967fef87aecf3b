"""
Multi-GPU layer: rays are independent for the whole trace, so a job shards by contiguous ray-index range over the
ranks (one process per GPU) with no communication during the trace; the only exchange is the all-reduce of reduced
products (pupil / PSF grid, statistics) afterwards -- ``Reducer.allreduce`` (NCCL over NVLink on GPUs).

The helpers here are backend-agnostic (they also run under ``gloo`` on CPU, which is how the CPU test-suite covers
the N > 1 logic).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """
    Contiguous range [first, first + count) of rank ``rank``: the first ``n_items % world`` ranks get one extra item,
    so the shards cover [0, n_items) exactly once and differ by at most one item.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_items), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def bind_to_device_cpus(device: int) -> list[int] | None:
    """
    Pin this process to the CPUs that are local to GPU ``device`` (NVML's ideal affinity: its NUMA node), so that the
    page-locked host buffers it allocates next -- and the threads that fill them -- sit next to that GPU's PCIe root.
    With one process per GPU on a two-socket node this keeps host<->device traffic off the socket interconnect.
    Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then).
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            token = visible.split(",")[device].strip()
            handle = (pynvml.nvmlDeviceGetHandleByUUID(token) if token.startswith(("GPU-", "MIG-"))
                      else pynvml.nvmlDeviceGetHandleByIndex(int(token)))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def allreduce_stats(stats, dist=None):
    """
    Combine RTB_N_STATS vectors over the ranks of the default process group: entries 0-7 are sums, 8/10 minima,
    9/11 maxima (include/rtb.h).  ``stats`` is a float64 torch tensor on the backend's device, (12,) or, for the
    buckets of a sweep, (n, 12); modified in place.
    """
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    sums = stats[..., 0:8].contiguous()
    mins = stats[..., [8, 10]].contiguous()
    maxs = stats[..., [9, 11]].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    stats[..., 0:8] = sums
    stats[..., [8, 10]] = mins
    stats[..., [9, 11]] = maxs
    return stats


def allreduce_grid(grid, dist=None):
    """Sum a (3, G, G) pupil grid over the ranks, in place."""
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return grid
    dist.all_reduce(grid, op=dist.ReduceOp.SUM)
    return grid


class Comm:
    """
    The library's own communicator (rtb_comm_* in include/rtb.h: NCCL behind the C ABI, no torch types): what a binder
    without PyTorch would use, and what ``Reducer.allreduce(comm=...)`` uses here.  Two collectives per exchange: one
    all-reduce for the grid, one all-gather + merge kernel for the statistics of every bucket.
    """

    def __init__(self, n_ranks: int, rank: int, unique_id: bytes, device: int):
        import ctypes as C
        from . import _ffi
        if len(unique_id) != _ffi.RTB_COMM_ID_BYTES:
            raise ValueError(f"unique_id must be {_ffi.RTB_COMM_ID_BYTES} bytes")
        self._lib = _ffi.lib()
        self._handle = C.c_void_p()
        self.n_ranks, self.rank, self.device = n_ranks, rank, device
        _ffi.check(self._lib.rtb_comm_init(C.byref(self._handle), n_ranks, rank, unique_id, device))

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _ffi
        buf = C.create_string_buffer(_ffi.RTB_COMM_ID_BYTES)
        _ffi.check(_ffi.lib().rtb_comm_unique_id(buf, _ffi.RTB_COMM_ID_BYTES))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int):
        """bootstrap over an initialised torch.distributed process group (any backend): rank 0's id is broadcast"""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        on_gpu = dist.get_backend() == "nccl"
        buf = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{device}" if on_gpu else "cpu")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return cls(world, rank, bytes(buf.cpu().numpy().tobytes()), device)

    def size(self) -> int:
        from . import _ffi
        n = self._lib.rtb_comm_size(self._handle)
        if n < 0:
            _ffi.check(n)
        return n

    def allreduce_grid(self, grid, stream_ptr: int):
        from . import _ffi
        _ffi.check(self._lib.rtb_comm_allreduce_grid(self._handle, grid.data_ptr(), grid.numel(), stream_ptr))

    def allreduce_stats(self, stats, stream_ptr: int):
        from . import _ffi
        n_buckets = stats.numel() // _ffi.RTB_N_STATS
        _ffi.check(self._lib.rtb_comm_allreduce_stats(self._handle, stats.data_ptr(), n_buckets, stream_ptr))

    def close(self):
        if self._handle:
            self._lib.rtb_comm_destroy(self._handle)
            self._handle = None


def merge_stats_host(parts) -> np.ndarray:
    """NumPy version of the same merge (used to check the collective path)."""
    parts = np.asarray(parts, dtype=np.float64)
    out = np.empty(12)
    out[0:8] = parts[:, 0:8].sum(axis=0)
    out[[8, 10]] = parts[:, [8, 10]].min(axis=0)
    out[[9, 11]] = parts[:, [9, 11]].max(axis=0)
    return out
