"""Host-buffer path (rtb_trace_host, pinned buffers, keep='last') for a few staging chunk sizes."""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) == 1:
    for mib in (8, 16, 32, 64, 128):
        env = dict(os.environ, RTB_HOST_CHUNK_MIB=str(mib))
        subprocess.run([sys.executable, __file__, str(mib)], env=env, check=True)
    sys.exit(0)

sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from ray_trace_pb_b200 import _ffi, engine  # noqa: E402

system, materials = bench.relay_system()
src, _ = bench.beam_source(1 << 24)
n = src.n_rays
host_in = _ffi.pinned_empty((n, 8))
host_in[:] = src.generate().cpu().numpy()
for keep, slabs in (("last", 1), ("all", 21)):
    m = n if slabs == 1 else n // 16
    out = _ffi.pinned_empty((slabs, m, 8))
    for _ in range(2):
        engine.trace_host(system.surfaces, materials, host_in[:m], keep=keep, out=out)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        engine.trace_host(system.surfaces, materials, host_in[:m], keep=keep, out=out)
    dt = (time.perf_counter() - t0) / reps
    gb = m * 64 * (1 + slabs) / 1e9
    print(f"chunk {sys.argv[1]:>4s} MiB  keep={keep:4s} {m:9d} rays  {dt * 1e3:8.2f} ms  {m * 10 / dt / 1e9:6.2f} G ray*surf/s  "
          f"{gb / dt:6.1f} GB/s over PCIe")
