"""
Workload-level helpers built on the fused source -> trace -> reduction kernels: the analyses the reference's example
scripts do after ``ray_trace`` (spot diagrams, centroids, RMS radii, pupil-phase maps, PSFs), without ever
materialising the rays.  Everything numerical runs in the kernels of librtb.so; this module only loops over
sources / fields / wavelengths and collects the small reduced results.
"""
from __future__ import annotations

import numpy as np

from . import device as dev


def _materials(system, initial_material, final_material):
    return [initial_material] + list(system.materials) + [final_material]


class _Launcher:
    """Packs the prescription once per wavelength and spreads independent launches over a few streams, so the tail
    of one launch overlaps the head of the next (the launches of a sweep are small: a few million rays each)."""

    def __init__(self, system, mats, device, n_streams=4):
        import torch
        self.torch = torch
        self.system, self.mats, self.device = system, mats, device
        self.packed = {}
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
        self.turn = 0
        self.origin_stream = torch.cuda.current_stream(device)

    def launch(self, src, first, count, precision, red):
        key = src.wavelength
        if key not in self.packed:
            self.packed[key] = dev.prepare(self.system.surfaces, self.mats,
                                           [src.wavelength] if np.isfinite(src.wavelength) else None)
        stream = self.streams[self.turn % len(self.streams)]
        self.turn += 1
        stream.wait_stream(self.origin_stream)
        with self.torch.cuda.stream(stream):
            dev.trace_source(self.system.surfaces, self.mats, src, first=first, count=count, keep="none",
                             precision=precision, reducer=red, device=self.device, packed=self.packed[key])

    def join(self):
        for st in self.streams:
            self.origin_stream.wait_stream(st)


def spot_statistics(system, initial_material, final_material, sources, slab: int = -2, origin=None,
                    e1=(1.0, 0.0, 0.0), e2=(0.0, 1.0, 0.0), precision: str = "f64", device: int = 0,
                    chunk: int = 1 << 27):
    """
    Spot statistics of one or several ray bundles at slab ``slab`` (default: at the last surface, the scripts'
    ``rays[-2]``): for each :class:`~ray_trace_pb_b200.device.RaySource` a dict with ``count``, ``centroid``,
    ``rms_radius``, ``mean_phase``, ``rms_phase``, ``u_range``, ``v_range``.  ``origin`` defaults to the centre of the
    surface the slab belongs to.  No ray I/O: sources of equal size (a field or wavelength sweep) are traced by ONE
    fused generate-trace-reduce launch (``dev.trace_sources``, one reduction bucket per source); otherwise one launch
    per source chunk, spread over streams.
    """
    single = isinstance(sources, dev.RaySource)
    sources = [sources] if single else list(sources)
    n_slabs = 2 * len(system.surfaces) + 1
    slab = slab + n_slabs if slab < 0 else slab
    if origin is None:
        origin = system.surfaces[max(slab - 1, 0) // 2].center
    mats = _materials(system, initial_material, final_material)
    if len(sources) > 1 and len({src.n_rays for src in sources}) == 1 and sources[0].n_rays <= chunk:
        red = dev.Reducer(slab, origin=origin, e1=e1, e2=e2, device=device, buckets=len(sources))
        dev.trace_sources(system.surfaces, mats, sources, keep="none", precision=precision, reducer=red, device=device)
        return red.stats()
    run = _Launcher(system, mats, device)
    reducers = []
    for src in sources:
        red = dev.Reducer(slab, origin=origin, e1=e1, e2=e2, device=device)     # reset on the caller's stream
        reducers.append(red)
        for first in range(0, src.n_rays, chunk):
            run.launch(src, first, min(chunk, src.n_rays - first), precision, red)
    run.join()
    results = [red.stats() for red in reducers]
    return results[0] if single else results


def reference_phase(system, mats, source, slab: int, device: int = 0, precision: str = "f64") -> float:
    """
    The phase of the bundle's central ray at ``slab`` (the ray in the middle of the source's index space: the chief ray of
    a fan / collimated / grid source with odd counts), or of the first ray of a 65-ray sample that reaches the slab;
    0.0 if none does.  Costs one small launch and one device->host read.
    """
    n = source.n_rays
    picks = sorted({n // 2, *np.linspace(0, n - 1, min(n, 64)).astype(np.int64).tolist()})
    phases = [float(dev.trace_source(system.surfaces, mats, source, first=int(i), count=1, keep=[slab],
                                     precision=precision, device=device)[0, 0, 6]) for i in [n // 2]]
    if not np.isfinite(phases[0]):
        phases = [float(dev.trace_source(system.surfaces, mats, source, first=int(i), count=1, keep=[slab],
                                         precision=precision, device=device)[0, 0, 6]) for i in picks]
    good = [p for p in phases if np.isfinite(p)]
    return good[0] if good else 0.0


def pupil_grid(system, initial_material, final_material, source, slab: int, origin, e1, e2, grid_n: int,
               half_width: float, phase_ref: float | None = None, precision: str = "f64", device: int = 0,
               chunk: int = 1 << 27):
    """
    Accumulate sum cos / sum sin / count of the phase of ``source``'s rays at slab ``slab`` on a ``grid_n`` x ``grid_n``
    grid spanning ``[-half_width, half_width)`` in the plane basis ``(origin, e1, e2)``.  Returns the
    :class:`~ray_trace_pb_b200.device.Reducer` (``.grid``, ``.stats()``, ``.psf()``, ``.allreduce()``).

    ``phase_ref`` is subtracted from every ray's phase before cos / sin are taken (a global phase: the PSF does not
    change).  The default, None, takes the central ray's phase (:func:`reference_phase`): accumulated phases are
    2 pi / wavelength x optical path -- 1e7 rad for millimetre paths at 532 nm -- and only their small differences
    matter; referencing them keeps the kernel's sine / cosine arguments small (no large-argument range reduction) and
    makes the statistics' phase sums (RMS wavefront error) well conditioned.  Pass 0.0 for absolute phases.
    """
    n_slabs = 2 * len(system.surfaces) + 1
    slab = slab + n_slabs if slab < 0 else slab
    mats = _materials(system, initial_material, final_material)
    if phase_ref is None:
        phase_ref = reference_phase(system, mats, source, slab, device=device, precision=precision)
    red = dev.Reducer(slab, origin=origin, e1=e1, e2=e2, grid_n=grid_n, half_width=half_width, phase_ref=phase_ref,
                      device=device)
    run = _Launcher(system, mats, device, n_streams=2)
    for first in range(0, source.n_rays, chunk):
        run.launch(source, first, min(chunk, source.n_rays - first), precision, red)
    run.join()
    return red


def axial_crossing(system, initial_material, final_material, wavelength: float, height: float,
                   pt=(0.0, 0.0, 0.0), normal=(0.0, 0.0, 1.0), device: int = 0):
    """
    Where the two meridional rays launched parallel to ``normal`` at ``+-height`` (along the bundle's first
    transverse axis) cross the centre ray behind the system (mean of the two crossings): the quantity behind the
    scripts' longitudinal-spherical-aberration and chromatic-focal-shift curves (``intersect_rays(axis_ray,
    rays[-1])``, e.g. scripts/2022_08_04_ACT508-100-B.py:158).  Returns a 3-vector (NaN where the rays do not meet
    to within the reference's 1e-12, as there).
    """
    from .raytrace import get_collimated_rays, intersect_rays
    probe = get_collimated_rays(pt, height, 3, wavelength, normal=normal)       # offsets -h, 0, +h in one plane
    traced = system.ray_trace(probe, initial_material, final_material, keep="last", device=device)[0]
    pts = intersect_rays(traced[1], traced[[0, 2]])
    return np.nanmean(pts, axis=0)
