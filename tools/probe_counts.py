"""What the lean kernel's probe launch finds on the bench workload: rays that reached each surface / lean-step failures."""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from ray_trace_pb_b200 import _ffi, device as dev  # noqa: E402

L = _ffi.lib()
L.rtb_tune(b"lean_min_rays", 0)
L.rtb_tune(b"keep_probe_counts", 1)
system, materials = bench.relay_system()
source, side = bench.beam_source(int(float(sys.argv[1]) if len(sys.argv) > 1 else 4e6))
rays = source.generate()
out = dev.trace_tensor(system.surfaces, materials, rays, keep="last", wavelengths=[bench.WAVELENGTH])
torch.cuda.synchronize()
n = len(system.surfaces)
buf = (ctypes.c_uint32 * (2 * n))()
L.rtb_last_probe_counts(buf, n)
print("surface: reached failed")
for k in range(n):
    print(f"  {k}: {buf[2 * k]} {buf[2 * k + 1]}")
