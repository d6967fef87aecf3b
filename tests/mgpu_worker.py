"""
Worker of tests/test_gpu_multi.py (one process per GPU under torchrun): each rank traces its contiguous shard of one
bundle with the fused reductions, the ranks exchange grid and statistics (a) through the library's own communicator
(rtb_comm_*: NCCL behind the C ABI) and (b) through torch.distributed, and every rank checks the result against the
single-GPU reduction of the WHOLE bundle done on its own device: counts exact, sums to 1e-10, min / max exact.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import systems  # noqa: E402
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402
from ray_trace_pb_b200 import _ffi, device as dev  # noqa: E402
from ray_trace_pb_b200.sharding import Comm, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert _ffi.lib().rtb_comm_available() > 0
    comm = Comm.from_torch_distributed(local)
    assert comm.size() == world

    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    src = dev.RaySource.grid([0, 0, 0], 14.0, 700, 0.785, normal=(np.sin(0.004), 0, np.cos(0.004)))
    n_buckets = 1
    whole = dev.Reducer(12, origin=(8.0, 0, 0), grid_n=256, half_width=9.0, device=local)
    dev.trace_source(system.surfaces, mats, src, keep="none", reducer=whole, device=local)

    first, count = shard_range(src.n_rays, rank, world)
    for route in ("rtb_comm", "torch"):
        part = dev.Reducer(12, origin=(8.0, 0, 0), grid_n=256, half_width=9.0, device=local)
        dev.trace_source(system.surfaces, mats, src, first=first, count=count, keep="none", reducer=part, device=local)
        part.allreduce(comm=comm if route == "rtb_comm" else None)
        torch.cuda.synchronize()
        got_s, want_s = part.stats_t.cpu().numpy(), whole.stats_t.cpu().numpy()
        assert got_s[0] == want_s[0] > 400_000, (route, got_s[0], want_s[0])
        np.testing.assert_allclose(got_s[1:8], want_s[1:8], rtol=1e-10, atol=1e-6, err_msg=route)
        np.testing.assert_array_equal(got_s[8:], want_s[8:], err_msg=route)
        got_g, want_g = part.grid_t.cpu().numpy(), whole.grid_t.cpu().numpy()
        np.testing.assert_array_equal(got_g[2], want_g[2], err_msg=route)
        np.testing.assert_allclose(got_g[:2], want_g[:2], rtol=0, atol=1e-9 * want_g[2].max(), err_msg=route)

    # the buckets of a sweep through the all-gather + merge kernel
    thetas = (0.0, 0.003, 0.006)
    sources = [dev.RaySource.grid([0, 0, 0], 12.0, 300, 0.785, normal=(np.sin(t), 0, np.cos(t))) for t in thetas]
    whole = dev.Reducer(19, buckets=3, device=local)
    dev.trace_sources(system.surfaces, mats, sources, keep="none", reducer=whole, device=local)
    first, count = shard_range(sources[0].n_rays, rank, world)
    part = dev.Reducer(19, buckets=3, device=local)
    dev.trace_sources(system.surfaces, mats, sources, first=first, count=count, keep="none", reducer=part, device=local)
    part.allreduce(comm=comm)
    torch.cuda.synchronize()
    got_s, want_s = part.stats_t.cpu().numpy(), whole.stats_t.cpu().numpy()
    np.testing.assert_array_equal(got_s[:, 0], want_s[:, 0])
    np.testing.assert_allclose(got_s[:, 1:8], want_s[:, 1:8], rtol=1e-10, atol=1e-6)
    np.testing.assert_array_equal(got_s[:, 8:], want_s[:, 8:])

    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"MGPU OK: {world} ranks, n_buckets {n_buckets}, count {int(want_s[0, 0])}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
