"""Wall time of the drop-in call System.ray_trace(numpy rays) -> numpy history (pageable host memory)."""
import sys
import time
from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import systems  # noqa: E402
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402

system = systems.relay10_system(rt, rtm)
vac = rtm.Vacuum()
for n_side in (100, 316, 1000, 2000):
    rays = systems.lattice_rays(n_side, 12.0, 0.0, 0.785)
    n = rays.shape[0]
    for keep in ("all", "last"):
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            out = system.ray_trace(rays, vac, vac, keep=keep)
            dt = time.perf_counter() - t0
            if rep:
                best = min(best, dt)
            del out
        print(f"N = {n:8d}  keep={keep:4s}  {best * 1e3:9.2f} ms   {n * 10 / best / 1e6:9.1f} M ray*surf/s "
              f"(reference NumPy: ~0.7 M/s per core)")
