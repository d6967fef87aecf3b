"""
`-m gpu`, needs >= 2 GPUs (skipped otherwise): the NCCL exchange step of the path -- an N-rank reduced grid / statistics
equals the single-GPU reduction of the whole bundle -- through the C ABI's rtb_comm_* and through torch.distributed.
One process per GPU under torchrun (tests/mgpu_worker.py).  The CPU suite covers the same host logic over gloo
(tests/test_sharding_gloo.py).
"""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_single_rank_communicator():
    """init / size / collectives / destroy with one rank (what a 1-GPU box can exercise)"""
    import torch
    from ray_trace_pb_b200 import _ffi, device as dev
    from ray_trace_pb_b200.sharding import Comm
    assert _ffi.lib().rtb_comm_available() > 0, "libnccl.so.2 could not be loaded"
    comm = Comm(1, 0, Comm.unique_id(), 0)
    assert comm.size() == 1
    red = dev.Reducer(0, grid_n=16, half_width=1.0)
    red.grid_t.fill_(2.0)
    red.stats_t.fill_(3.0)
    red.allreduce(comm=comm)
    torch.cuda.synchronize()
    assert float(red.grid_t.sum()) == 2.0 * 3 * 16 * 16 and float(red.stats_t.sum()) == 36.0
    comm.close()


@pytest.mark.parametrize("n_ranks", [2, 4])
def test_n_rank_reduction_equals_single_gpu(n_ranks):
    if _n_gpus() < n_ranks:
        pytest.skip(f"needs {n_ranks} GPUs, this box has {_n_gpus()}")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + n_ranks),
                        str(ROOT / "tests" / "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
