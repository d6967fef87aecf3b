#!/usr/bin/env python
"""
The five BASELINE.json configurations at full size on one B200 (config 5 sharded over more GPUs is bench.py).

    python examples/run_configs.py [--configs 1 2 3 4 5] [--precision f64]

Each block prints what the corresponding reference script plots / reports, plus the wall time of the GPU part.
Nothing here touches the reference or the CPU oracle.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import systems  # noqa: E402  (workload definitions shared with the tests)
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402
from ray_trace_pb_b200 import analysis, device as dev  # noqa: E402


def timed(fn):
    """Second of two runs (the first pays one-off costs: CUDA module loading of the kernel variant, stream creation)."""
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def config1(precision):
    """scripts/2022_10_27_plano_convex_lens.py scaled up: 1001 x 1000 collimated rays, OPL vs the analytic formula"""
    system, m_in, m_out, _ = systems.plano_convex(rt, rtm)
    mats = [m_in] + system.materials + [m_out]
    src = dev.RaySource.collimated([0, 0, -5], 25.4, 1001, 0.5, nphis=1000)
    (rays, last), dt = timed(lambda: (src.generate(), dev.trace_source(system.surfaces, mats, src, keep="last",
                                                                        precision=precision)[0]))
    h = torch.hypot(rays[:, 0], rays[:, 1]).cpu().numpy()
    opl = (last[:, 6] / (2 * np.pi / 0.5)).cpu().numpy()
    n, t0, t1, R, dz = 1.3, 2.679486355, 1, 100, 5
    with np.errstate(all="ignore"):
        sag = R - np.sqrt(R**2 - h**2)
        ref = dz + n * t0 + n * t1 - n * sag + sag / (np.sqrt(1 - n**2 * h**2 / R**2) * np.sqrt(R**2 - h**2) / R + n * h**2 / R**2)
    ok = np.isfinite(opl)
    print(f"config 1: {src.n_rays} rays x 3 surfaces in {dt * 1e3:.1f} ms; {ok.sum()} valid; "
          f"max |OPL - analytic| = {np.abs(opl[ok] - ref[ok]).max():.2e} mm")


def config2(precision):
    """scripts/2022_08_04_ACT508-100-B.py: AC508-100-B, 3 wavelengths x 4096^2 pupil grid: spots + chromatic shift"""
    doublet = rt.Doublet(rtm.Nlak22(), rtm.Nsf6ht(), radius_crown=65.8, radius_flint=-280.6, radius_interface=-56,
                         thickness_crown=13.0, thickness_flint=2.0, aperture_radius=25.4, names="AC508-100-B")
    vac = rtm.Vacuum()
    wls = (0.7065, 0.855, 1.015)
    cps = {w: doublet.get_cardinal_points(w, vac, vac) for w in wls}
    for w in wls:
        print(f"config 2: {w} um paraxial efl = {cps[w][7]:.3f} mm, bfl = {cps[w][1][2] - 15.0:.3f} mm (catalog 100 / 91.5)")
    focus_z = float(cps[0.855][1][2])
    system = rt.System([rt.FlatSurface([0, 0, -5.0], [0, 0, 1], 25.4)], []).concatenate(doublet, vac)
    system = system.concatenate(rt.FlatSurface([0, 0, focus_z], [0, 0, 1], 25.4), vac)
    sources = [dev.RaySource.grid([0, 0, -10.0], 10.0, 4096, w) for w in wls]
    stats, dt = timed(lambda: analysis.spot_statistics(system, vac, vac, sources, slab=-2, precision=precision))
    for w, s in zip(wls, stats):
        print(f"config 2: {w} um  {s['count']} rays at the d-line focal plane: centroid ({s['centroid'][0]:+.2e}, "
              f"{s['centroid'][1]:+.2e}) mm, RMS spot radius {s['rms_radius'] * 1e3:.3f} um")
    print(f"config 2: 3 x {sources[0].n_rays} rays x 5 surfaces in {dt * 1e3:.1f} ms "
          f"({3 * sources[0].n_rays * 5 / dt / 1e9:.1f} G ray*surf/s, fused source + trace + statistics)")
    shifts = [analysis.axial_crossing(system, vac, vac, w, 1e-3, pt=(0, 0, -10.0))[2] for w in wls]
    print("config 2: paraxial-ring focus z = " + ", ".join(f"{z:.4f}" for z in shifts) +
          f" mm -> chromatic focal shift {shifts[-1] - shifts[0]:+.4f} mm")


def config3(precision):
    """scripts/2022_08_24_relay_astigmatism.py: 10-surface relay, 32 field bundles x 2048^2 rays, spot per field"""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    thetas = np.linspace(0, np.pi / 180, 32)
    sources = []
    for th in thetas:
        nrm = np.array([np.sin(th), 0, np.cos(th)])
        sources.append(dev.RaySource.grid([0, 0, 0], 12.0, 2048, 0.785, normal=nrm / np.linalg.norm(nrm)))
    stats, dt = timed(lambda: analysis.spot_statistics(system, vac, vac, sources, slab=-2, precision=precision))
    for k in (0, 10, 21, 31):
        s = stats[k]
        print(f"config 3: field {np.degrees(thetas[k]):.3f} deg: {s['count']} valid, centroid x = {s['centroid'][0]:+.4f} mm, "
              f"RMS spot {s['rms_radius'] * 1e3:.2f} um, RMS phase {s['rms_phase']:.3f} rad")
    n = sum(src.n_rays for src in sources)
    print(f"config 3: 32 x {sources[0].n_rays} rays x 10 surfaces in {dt * 1e3:.1f} ms ({n * 10 / dt / 1e9:.1f} G ray*surf/s)")
    rays = np.concatenate((rt.get_collimated_rays([0, 0, 0], 1.0, 19, 0.785),
                           rt.get_collimated_rays([0, 0, 0], 1.0, 19, 0.785, phi_start=np.pi / 2)))
    hist = system.ray_trace(rays, vac, vac)
    fm = rt.intersect_rays(hist[-2, 8], hist[-2, 10])
    fs = rt.intersect_rays(hist[-2, 19 + 8], hist[-2, 19 + 10])
    print(f"config 3: meridional - sagittal focus (lenses 5 mm off axis) = {fm[0, 2] - fs[0, 2]:+.5f} mm")


def config4(precision):
    """scripts/2022_01_25_ray_trace_ideal_opm.py: ideal OPM, 16001 x 16000 ray fan -> O3 pupil phase -> PSF"""
    system, m_in, m_out, alpha1, theta = systems.opm_system(rt, rtm)
    src = dev.RaySource.fan([1e-3, 1e-3, 1e-3 * np.tan(theta)], alpha1, 16001, 532e-6, nphis=16000)
    o3n = system.surfaces[8].normal
    e2 = np.array([0.0, 1.0, 0.0])
    e1 = np.cross(e2, o3n)
    chief = system.ray_trace(rt.get_ray_fan([1e-3, 1e-3, 1e-3 * np.tan(theta)], 0.0, 1, 532e-6), m_in, m_out)
    red, dt = timed(lambda: analysis.pupil_grid(system, m_in, m_out, src, slab=-5, origin=system.surfaces[8].center,
                                                e1=e1, e2=e2, grid_n=2048, half_width=3.2,
                                                phase_ref=float(chief[-5, 0, 6]), precision=precision))
    s = red.stats()
    print(f"config 4: {src.n_rays} rays x 11 surfaces -> 2048^2 pupil grid in {dt * 1e3:.1f} ms "
          f"({src.n_rays * 11 / dt / 1e9:.1f} G ray*surf/s); {s['count']} rays reach the O3 pupil, "
          f"footprint u {s['u_range'][0]:+.3f}..{s['u_range'][1]:+.3f}, v {s['v_range'][0]:+.3f}..{s['v_range'][1]:+.3f} mm")
    psf, dt2 = timed(lambda: red.psf(257, 1.0 / (532e-6 * 200.0) * 2e-3, normalize_by_count=True))
    psf = (psf / psf.max()).cpu().numpy()
    pv, pu = np.unravel_index(int(psf.argmax()), psf.shape)
    box = psf[max(pv - 8, 0):pv + 9, max(pu - 8, 0):pu + 9]
    print(f"config 4: 257^2 PSF samples (2 um pitch at the f = 200 mm tube lens) in {dt2 * 1e3:.2f} ms; peak at "
          f"({(pu - 128) * 2:+d}, {(pv - 128) * 2:+d}) um from the axis (the image of the 1 um off-axis source at "
          f"~140x), energy within +-16 um of the peak = {float(box.sum() / psf.sum()):.3f}")


def config5(precision):
    """scripts/2024_08_08_achromat_imaging.py: 9-surface achromat 4f system, 1.25e8 rays (one GPU's share of 1e9)"""
    system = systems.achromat_imaging_system(rt, rtm)
    vac = rtm.Vacuum()
    src = dev.RaySource.fan([2.0, 0, 0], 4 * np.pi / 180, 11181, 0.635, nphis=11180)
    pupil = system.surfaces[4]
    red, dt = timed(lambda: analysis.pupil_grid(system, vac, vac, src, slab=2 * 4 + 2, origin=pupil.center,
                                                e1=(1, 0, 0), e2=(0, 1, 0), grid_n=2048, half_width=8.0,
                                                precision=precision))
    s = red.stats()
    print(f"config 5: {src.n_rays} rays x 9 surfaces (Ebaf11/Nsf11 doublets, host-tabulated n) -> 2048^2 pupil grid in "
          f"{dt * 1e3:.1f} ms ({src.n_rays * 9 / dt / 1e9:.1f} G ray*surf/s); {s['count']} rays through the stop")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, nargs="*", default=[1, 2, 3, 4, 5])
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--repeat", type=int, default=1, help="run every config this many times (from the second run on the "
                    "lean kernel's verdict cache picks the pure instantiations)")
    args = ap.parse_args()
    print(f"device: {torch.cuda.get_device_name(0)}; precision {args.precision}")
    for c in args.configs:
        for _ in range(args.repeat):
            {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}[c](args.precision)
