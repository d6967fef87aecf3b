"""Probe counts of the lean kernel on the OPM fan (config 4)."""
import ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import systems
import ray_trace_pb_b200.materials as rtm
import ray_trace_pb_b200.raytrace as rt
from ray_trace_pb_b200 import _ffi, device as dev
L = _ffi.lib()
L.rtb_tune(b"lean_min_rays", 0); L.rtb_tune(b"keep_probe_counts", 1)
system, m_in, m_out, alpha1, theta = systems.opm_system(rt, rtm)
mats = [m_in] + system.materials + [m_out]
src = dev.RaySource.fan([1e-3, 1e-3, 1e-3 * np.tan(theta)], alpha1, 2001, 532e-6, nphis=2000)
out = dev.trace_source(system.surfaces, mats, src, keep="last")
torch.cuda.synchronize()
n = len(system.surfaces)
buf = (ctypes.c_uint32 * (2 * n))()
L.rtb_last_probe_counts(buf, n)
for k in range(n):
    print(k, type(system.surfaces[k]).__name__, buf[2 * k], buf[2 * k + 1])
print("alive at end", int(torch.isfinite(out[0, :, 0]).sum()), "of", src.n_rays)
