"""Fuzz of the lean kernels (csrc/trace_lean.cu): random lens trains the lean kernel takes -- on-axis spheres (centred and
decentred), on-axis and tilted flats, perfect lenses on and off the z axis -- under bundles of every character (skew,
collimated along z, meridional, launched ON the first flat, through a lens's focal point, with NaN / inf rows), final
slab and fused reductions, through
    the round-1 kernels (lean off),
    the probe-driven lean kernels           (rtb_tune lean_pure = 0),
    the pure kernels with no verdict        (lean_pure = 2: whatever the plain lean steps cannot take goes through redo_ray),
    the verdict cache                       (lean_pure = 1, the system traced three times: probe-driven, then pure),
each against the CPU oracle, bit for bit.   python tools/fuzz_lean_systems.py [first_seed] [n_seeds]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parity  # noqa: E402
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402
from oracle import oracle  # noqa: E402  (checker only)
from ray_trace_pb_b200 import _ffi, device as dev  # noqa: E402

L = _ffi.lib()


def unit(v):
    v = np.asarray(v, dtype=float)
    return v / np.sqrt((v * v).sum())


def lean_system(seed):
    rng = np.random.default_rng(50_000 + seed)
    glasses = [rtm.Vacuum, lambda: rtm.Constant(1.33), rtm.Bk7, rtm.Sf10, rtm.Nlak22, rtm.FusedSilica, rtm.Nsf6]
    n_surf = int(rng.integers(3, 13))
    kinds_pool = [["sphere", "sphere", "sphere", "flat"],                        # KINDS 0 / 2: refracting only
                  ["sphere", "sphere", "flat", "tilted", "lens", "lens"]][int(rng.integers(0, 2))]
    surfaces, z = [], 0.0
    for k in range(n_surf):
        kind = "flat" if (k == 0 and rng.random() < 0.5) else str(rng.choice(kinds_pool))
        aperture = float(rng.uniform(6.0, 15.0))
        off = rng.uniform(-1.0, 1.0, 2) * (rng.random() < 0.3)
        if kind == "sphere":
            radius = float(rng.uniform(25.0, 250.0) * (1 if rng.random() < 0.5 else -1))
            surfaces.append(rt.SphericalSurface(radius, [off[0], off[1], z + radius], aperture))
        elif kind == "flat":
            surfaces.append(rt.FlatSurface([off[0], off[1], z], [0.0, 0.0, 1.0], aperture))
        elif kind == "tilted":
            t = rng.uniform(-0.1, 0.1, 2)
            surfaces.append(rt.FlatSurface([off[0], off[1], z], unit([t[0], t[1], 1.0]), aperture))
        else:
            t = rng.uniform(-0.05, 0.05, 2) * (rng.random() < 0.3)
            normal = unit([t[0], t[1], 1.0]) if t.any() else np.array([0.0, 0.0, 1.0])
            surfaces.append(rt.PerfectLens(float(rng.uniform(30.0, 120.0)), [off[0], off[1], z], normal,
                                           float(rng.uniform(0.2, 0.6))))
        z += float(rng.uniform(2.0, 25.0))
    pick = lambda: glasses[int(rng.integers(0, len(glasses)))]()
    system = rt.System(surfaces, [pick() for _ in range(n_surf - 1)])
    return system, pick(), pick(), rng


def bundle(rng, which, n, system):
    wl = rng.uniform(0.45, 1.0, 2)
    rays = np.zeros((n, 8))
    r = 7.0 * np.sqrt(rng.random(n))
    phi = rng.uniform(0, 2 * np.pi, n)
    rays[:, 0], rays[:, 1], rays[:, 2] = r * np.cos(phi), r * np.sin(phi), -8.0
    d = rng.standard_normal((n, 3)) * np.array([0.04, 0.04, 0.0]) + np.array([0.0, 0.0, 1.0])
    rays[:, 6] = rng.uniform(0, 40, n)
    rays[:, 7] = rng.choice(wl, size=n)
    if which == "collimated":
        d[:] = [0.0, 0.0, 1.0]
    elif which == "meridional":
        rays[:, 1] = 0.0
        d[:, 1] = 0.0
    elif which == "on_first":
        rays[:, 2] = float(system.surfaces[0].center[2])          # t = +-0 at a first flat at this z
        rays[::2, 2] = -0.0 + rays[::2, 2]
    elif which == "dirty":
        rays[rng.integers(0, n, n // 50)] = np.nan
        rays[rng.integers(0, n, n // 50), rng.integers(0, 7, n // 50)] = np.inf
        rays[rng.integers(0, n, n // 50), 7] = rng.uniform(0.4, 1.0, n // 50)      # unlisted wavelengths
    rays[:, 3:6] = d / np.sqrt((d * d).sum(axis=1, keepdims=True))
    return rays


first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n_rays = 6000
bad = 0
pure_before = L.rtb_pure_launch_count()
traces = 0
_ffi.check(L.rtb_tune(b"lean_min_rays", 0))
for seed in range(first, first + count):
    system, m_in, m_out, rng = lean_system(seed)
    mats = [m_in] + list(system.materials) + [m_out]
    n_slabs = 2 * len(system.surfaces) + 1
    for which in ("skew", "collimated", "meridional", "on_first", "dirty"):
        rays = bundle(rng, which, n_rays, system)
        hist = oracle.ray_trace(system, rays, m_in, m_out, n_threads=8)
        want = parity.canonical(hist[[-1]])
        d_rays = torch.from_numpy(rays).cuda()
        # a reduction slab the lean kernel takes: after a surface, or at a refracting one
        slab = int(rng.integers(1, n_slabs))
        if slab % 2 == 1 and isinstance(system.surfaces[(slab - 1) // 2], rt.PerfectLens):
            slab += 1
        origin = tuple(np.nanmean(hist[slab][:, 0:3], axis=0)) if np.isfinite(hist[slab][:, 0]).any() else (0, 0, 0)
        ref_stats = oracle.reduce_stats(hist[slab], origin, (1, 0, 0), (0, 1, 0))
        ref_grid = oracle.reduce_grid(hist[slab], origin, (1, 0, 0), (0, 1, 0), 32, 10.0)
        for mode, reps in (("general", 1), (0, 1), (2, 1), (1, 3)):
            _ffi.check(L.rtb_tune(b"lean_min_rays", -1 if mode == "general" else 0))
            if mode != "general":
                _ffi.check(L.rtb_tune(b"lean_pure", mode))
            for rep in range(reps):
                red = dev.Reducer(slab, origin=origin, grid_n=32, half_width=10.0)
                wls = sorted(set(rays[np.isfinite(rays[:, 7]), 7].tolist()))[:8]
                out = dev.trace_tensor(system.surfaces, mats, d_rays, keep="last", wavelengths=wls, reducer=red)
                torch.cuda.synchronize()
                traces += 1
                got = parity.canonical(out.cpu().numpy())
                stats = red.stats_t.cpu().numpy()
                grid = red.grid.cpu().numpy()
                ok = np.array_equal(got, want) and stats[0] == ref_stats[0] and np.array_equal(stats[8:], ref_stats[8:]) \
                    and np.allclose(stats[1:8], ref_stats[1:8], rtol=1e-10, atol=1e-6) \
                    and np.array_equal(grid[2], ref_grid[2]) \
                    and np.allclose(grid[:2], ref_grid[:2], rtol=0, atol=1e-9 * max(1.0, ref_grid[2].max()))
                if not ok:
                    bad += 1
                    print(f"seed {seed} bundle {which} mode {mode} rep {rep} slab {slab}: MISMATCH "
                          f"(final slab equal: {np.array_equal(got, want)}, count {stats[0]} vs {ref_stats[0]})")
print(f"seeds {first}..{first + count - 1}: {traces} traces, {bad} mismatches; "
      f"{L.rtb_pure_launch_count() - pure_before} of them ran the pure kernels")
sys.exit(1 if bad else 0)
