"""One sweep launch (rtb_trace_sources) against one launch per source: device time by CUDA events."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import systems  # noqa: E402
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402
from ray_trace_pb_b200 import device as dev  # noqa: E402

system = systems.relay10_system(rt, rtm)
vac = rtm.Vacuum()
mats = [vac] + list(system.materials) + [vac]
n_src, side = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
sources = []
for th in np.linspace(0, np.pi / 180, n_src):
    nrm = np.array([np.sin(th), 0, np.cos(th)])
    sources.append(dev.RaySource.grid([0, 0, 0], 12.0, side, 0.785, normal=nrm / np.linalg.norm(nrm)))
packed = dev.prepare(system.surfaces, mats, [0.785])
total = n_src * sources[0].n_rays * 10


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


red = dev.Reducer(20, buckets=n_src)
ms = timed(lambda: dev.trace_sources(system.surfaces, mats, sources, keep="none", reducer=red, packed=packed))
print(f"sweep launch     : {ms:8.3f} ms  {total / ms / 1e6:7.2f} G ray*surf/s")
reds = [dev.Reducer(20) for _ in sources]
ms = timed(lambda: [dev.trace_source(system.surfaces, mats, s, keep="none", reducer=r, packed=packed)
                    for s, r in zip(sources, reds)])
print(f"launch per source: {ms:8.3f} ms  {total / ms / 1e6:7.2f} G ray*surf/s")
a = np.stack([r.stats()["raw"] for r in reds])
st = red.stats()
b = np.stack([s["raw"] for s in (st if isinstance(st, list) else [st])])
print("counts equal:", np.array_equal(a[:, 0], b[:, 0]))
