"""
CPU test (no GPU): the ctypes stub of INTEGRATION.md section 2 (tests/reference_stub.py) run against the REAL reference
classes -- the unmodified package in baseline/_ref/ (installed by baseline/install_ref.sh in the build container and
shipped with the snapshot; /root/reference/src is the fallback there) -- up to, and not including, the device call:
every struct the stub packs from the reference's own objects must equal, byte for byte, what this framework's host
layer packs from its mirror of the same system, and the refractive-index tables must be the same bits.
Skipped where the reference is not installed.
"""
import ctypes
import sys
import types
from pathlib import Path

import numpy as np
import pytest

import systems

ROOT = Path(__file__).resolve().parent.parent


def _reference_modules():
    for base in (ROOT / "baseline" / "_ref", Path("/root/reference/src")):
        if (base / "raytrace" / "raytrace.py").exists():
            break
    else:
        pytest.skip("the reference is not installed (baseline/install_ref.sh)")
    for name in ("matplotlib", "matplotlib.figure", "matplotlib.axes", "matplotlib.axes._axes", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.figure"].Figure = object
    sys.modules["matplotlib.axes._axes"].Axes = object
    saved = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "raytrace" or k.startswith("raytrace.")]}
    sys.path.insert(0, str(base))
    try:
        import raytrace.materials as ref_rtm
        import raytrace.raytrace as ref_rt
        assert Path(ref_rt.__file__).resolve().is_relative_to(base.resolve())
    finally:
        sys.path.remove(str(base))
        for k in [k for k in sys.modules if k == "raytrace" or k.startswith("raytrace.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_rt, ref_rtm


@pytest.mark.parametrize("builder", ["relay10", "opm", "edge_mix", "doublet_nlak22", "mirrors", "reversed_doublet"])
def test_stub_packs_the_real_reference_classes(builder, rt, rtm):
    import reference_stub
    from ray_trace_pb_b200 import engine
    ref_rt, ref_rtm = _reference_modules()
    system, m_in, m_out, rays = getattr(systems, builder)(ref_rt, ref_rtm)            # the reference's own objects
    sysd, packed_rays, _keep = reference_stub.pack(system, rays, m_in, m_out)
    assert sysd.n_surfaces == len(system.surfaces) and packed_rays.shape == (len(rays), 8)

    mirror, mm_in, mm_out = systems.rebuild_system(systems.describe_system(system, m_in, m_out), rt, rtm)
    wl = np.unique(rays[~np.isnan(rays[:, 7]), 7])
    ours = engine.pack_system(mirror.surfaces, [mm_in] + mirror.materials + [mm_out], wl)
    assert ours.sys.n_surfaces == sysd.n_surfaces and ours.sys.n_wavelengths == sysd.n_wavelengths
    for k in range(sysd.n_surfaces):
        a, b = sysd.surfaces[k], ours.sys.surfaces[k]
        b.hints = 0                                          # (speed hints are this framework's own business)
        assert bytes(a) == bytes(b), f"surface {k} of {builder}"
    n_tab = (sysd.n_wavelengths + 1) * (sysd.n_surfaces + 1)
    ta = np.ctypeslib.as_array(sysd.n_table, shape=(n_tab,))
    tb = np.ctypeslib.as_array(ours.sys.n_table, shape=(n_tab,))
    assert np.array_equal(ta.view(np.uint64), tb.view(np.uint64)), "refractive-index table"
    assert ctypes.sizeof(reference_stub.RtbSurface) == ctypes.sizeof(type(ours.sys.surfaces[0]))
