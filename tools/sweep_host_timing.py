"""Wall-clock pieces of analysis.spot_statistics on a sweep (where the host time goes)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import systems
import ray_trace_pb_b200.materials as rtm
import ray_trace_pb_b200.raytrace as rt
from ray_trace_pb_b200 import analysis, device as dev

system = systems.relay10_system(rt, rtm)
vac = rtm.Vacuum()
mats = [vac] + list(system.materials) + [vac]
sources = []
for th in np.linspace(0, np.pi / 180, 32):
    nrm = np.array([np.sin(th), 0, np.cos(th)])
    sources.append(dev.RaySource.grid([0, 0, 0], 12.0, 2048, 0.785, normal=nrm / np.linalg.norm(nrm)))

def wall(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:40s} {best * 1e3:8.2f} ms")
    return out

wall("spot_statistics (sweep route)", lambda: analysis.spot_statistics(system, vac, vac, sources, slab=-2))
red = wall("Reducer(buckets=32)", lambda: dev.Reducer(20, buckets=32))
packed = wall("prepare", lambda: dev.prepare(system.surfaces, mats, [0.785]))
wall("trace_sources packed", lambda: dev.trace_sources(system.surfaces, mats, sources, keep="none", reducer=red, packed=packed))
wall("trace_sources unpacked", lambda: dev.trace_sources(system.surfaces, mats, sources, keep="none", reducer=red))
wall("stats()", lambda: red.stats())
wall("per-source loop packed", lambda: [dev.trace_source(system.surfaces, mats, s, keep="none", reducer=red2, packed=packed) for s, red2 in zip(sources, [dev.Reducer(20) for _ in sources])])
