#!/usr/bin/env python
"""
The reference's own NumPy path, timed on this box's host cores (SURVEY.md 8d "CPU baseline timing", north star: "the
reference's numpy path, timed on the box's host cores in the same run with the core count stated").

    python baseline/numpy_leg.py --rays 1000000 --procs 1  --reps 3
    python baseline/numpy_leg.py --rays 1000000 --procs 16 --reps 3

Runs the UNMODIFIED reference installed in baseline/_ref/ (baseline/install_ref.sh): `System.ray_trace` of
raytrace/raytrace.py:641-661 on the bench workload (the 10-surface relay of tests/systems.py: relay10_system, a
collimated Cartesian bundle at 0.785 um), full (2S+1, N, 8) history -- its only output mode.  With --procs P the
bundle is cut into P contiguous index ranges and P worker processes trace them at the same time (NumPy's elementwise
kernels are single-threaded); the time of a repetition is the wall clock from a common start barrier to the last
worker's finish.  Prints ONE JSON line.  Nothing of ray_trace_pb_b200 / librtb.so is imported here.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"


def import_reference():
    """matplotlib is not in the image and the hot path never touches it: empty stand-in modules."""
    for name in ("matplotlib", "matplotlib.figure", "matplotlib.axes", "matplotlib.axes._axes", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.figure"].Figure = object
    sys.modules["matplotlib.axes._axes"].Axes = object
    for k in [k for k in sys.modules if k == "raytrace" or k.startswith("raytrace.")]:
        del sys.modules[k]
    sys.path.insert(0, str(REF))
    import raytrace.materials as rtm
    import raytrace.raytrace as rt
    assert Path(rt.__file__).resolve().is_relative_to(REF.resolve()), rt.__file__
    sys.path.remove(str(REF))
    return rt, rtm


def _worker(rank, n_procs, n_side, reps, start, done, times):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    rt, rtm = import_reference()
    sys.path.insert(0, str(HERE.parent / "tests"))
    import systems                                      # workload definitions only (shared with the golden generator)
    system = systems.relay10_system(rt, rtm)
    rays = systems.lattice_rays(n_side, 12.0, 0.0, 0.785)
    n = rays.shape[0]
    lo, hi = rank * n // n_procs, (rank + 1) * n // n_procs
    mine = np.ascontiguousarray(rays[lo:hi])
    del rays
    vac = rtm.Vacuum()
    with np.errstate(all="ignore"):
        system.ray_trace(mine[: min(len(mine), 2000)], vac, vac)           # warm-up (imports, allocator)
    for r in range(reps):
        start.wait()
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            hist = system.ray_trace(mine, vac, vac)
        times[rank * reps + r] = time.perf_counter() - t0
        assert hist.shape == (2 * len(system.surfaces) + 1, hi - lo, 8)
        del hist
        done.wait()


def run(n_rays: int, n_procs: int, reps: int, budget_s: float) -> dict:
    n_side = int(round(n_rays ** 0.5))
    ctx = mp.get_context("fork")
    start, done = ctx.Barrier(n_procs + 1), ctx.Barrier(n_procs + 1)
    times = ctx.Array("d", n_procs * reps)
    procs = [ctx.Process(target=_worker, args=(r, n_procs, n_side, reps, start, done, times)) for r in range(n_procs)]
    for p in procs:
        p.start()
    walls = []
    t_begin = time.perf_counter()
    for r in range(reps):
        start.wait()
        t0 = time.perf_counter()
        done.wait()
        walls.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and r + 1 < reps:
            # over the time budget: let the workers run out their remaining repetitions on a released barrier
            for _ in range(r + 1, reps):
                start.wait()
                done.wait()
            break
    for p in procs:
        p.join()
    best = min(walls)
    surfaces = 10
    return {"value": n_side * n_side * surfaces / best, "unit": "ray*surfaces/s", "cores": n_procs, "kind": "reference",
            "impl": "QI2lab/ray_trace_pb System.ray_trace (NumPy), unmodified, from baseline/_ref",
            "sample": f"{n_side}x{n_side} = {n_side * n_side} rays x {surfaces} surfaces, full (21, N, 8) history, "
                      f"{n_procs} process(es), best of {len(walls)}",
            "seconds": walls}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=1e6)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--budget-s", type=float, default=1e9, help="stop repeating once this much time has gone by")
    args = ap.parse_args()
    if not (REF / "raytrace" / "raytrace.py").exists():
        print(json.dumps({"unavailable": "baseline/_ref is empty: run baseline/install_ref.sh in the build container"}))
        return
    print(json.dumps(run(int(args.rays), args.procs, args.reps, args.budget_s)))


if __name__ == "__main__":
    main()
