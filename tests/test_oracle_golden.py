"""
CPU tests (no GPU): the oracle against the golden vectors written from the reference itself.

These are the pins the GPU parity tests rest on: if oracle/rt_oracle.c drifts from the reference's arithmetic,
this file fails before any CUDA result is trusted.
"""

import numpy as np
import pytest

import parity
import systems
from conftest import load_golden


def _tables_match(g, materials):
    """True when this host's NumPy reproduces the golden refractive indices bit for bit (np.power platform check)."""
    uniq = g["unique_wavelengths"]
    with np.errstate(all="ignore"):
        ntab = np.stack([np.asarray(m.n(uniq), dtype=float).reshape(-1) for m in materials], axis=1)
    return np.array_equal(parity.canonical(ntab), parity.canonical(g["n_table"]))


@pytest.mark.parametrize("name", sorted(systems.CASES))
def test_oracle_reproduces_reference_history(name, rt, rtm, oracle):
    g = load_golden(name)
    system, m_in, m_out = systems.rebuild_system(g["system"], rt, rtm)
    got = oracle.ray_trace(system, g["rays_in"], m_in, m_out)
    materials = [m_in] + system.materials + [m_out]
    if name in systems.POWER_DEPENDENT and not _tables_match(g, materials):
        parity.assert_close_same_mask(got, g["history"], rtol=1e-10, what=name)
    else:
        parity.assert_bit_identical(got, g["history"], name)


def test_oracle_input_shapes(rt, rtm, oracle):
    g = load_golden("input_shapes")
    system, m_in, m_out = systems.rebuild_system(g["system"], rt, rtm)
    parity.assert_bit_identical(oracle.ray_trace(system, g["rays"][3], m_in, m_out), g["single"], "(8,) input")
    parity.assert_bit_identical(oracle.ray_trace(system, g["pre"], m_in, m_out), g["ext"], "(K,N,8) input")


BIG = {
    "relay10_lattice": lambda: systems.lattice_rays(320, 14.0, 0.0, 0.785, tilt=(0.004, -0.002)),
    "relay10_lattice_wide": lambda: systems.lattice_rays(200, 27.0, 0.0, 0.785, tilt=(0.0, 0.0), converge=1e-4),
    "plano_convex_lattice": lambda: systems.lattice_rays(256, 26.0, -5.0, 0.5),
    "opm_lattice": lambda: systems.lattice_rays(256, 1e-3, 1e-3, 532e-6, tilt=(0.0, 0.0), converge=600.0),
    "long_train_lattice": lambda: systems.lattice_rays(64, 15.5, 0.0, 0.5876, tilt=(0.002, -0.001)),
}


@pytest.mark.parametrize("name", sorted(BIG))
def test_oracle_big_batch_checksum(name, rt, rtm, oracle, checksums):
    """1e5-ray batches: sha256 of the reference's full history, recorded by make_golden.py."""
    ref = checksums["big"][name]
    system, m_in, m_out = systems.rebuild_system(ref["system"], rt, rtm)
    rays = BIG[name]()
    assert rays.shape[0] == ref["n_rays"]
    hist = oracle.ray_trace(system, rays, m_in, m_out, n_threads=4)
    assert int(np.isnan(hist[-1, :, 0]).sum()) == ref["nan_rays_at_end"]
    assert parity.digest(hist) == ref["sha256_history"]
    assert parity.digest(hist[-1]) == ref["sha256_last"]


@pytest.mark.parametrize("seed", range(systems.N_RANDOM_SYSTEMS))
def test_oracle_random_systems(seed, rt, rtm, oracle, checksums):
    """40 random systems (all surface kinds, decentred / tilted, random glasses): sha256 of the reference's history"""
    system, m_in, m_out, rays = systems.random_system(rt, rtm, seed)
    ref = checksums["random"][str(seed)]
    assert len(system.surfaces) == ref["n_surfaces"]
    hist = oracle.ray_trace(system, rays, m_in, m_out)
    assert int(np.isfinite(hist[-1, :, 0]).sum()) == ref["alive_at_end"]
    assert parity.digest(hist) == ref["sha256_history"]


def test_oracle_threads_do_not_change_results(rt, rtm, oracle):
    g = load_golden("edge_mix")
    system, m_in, m_out = systems.rebuild_system(g["system"], rt, rtm)
    a = oracle.ray_trace(system, g["rays_in"], m_in, m_out, n_threads=1)
    b = oracle.ray_trace(system, g["rays_in"], m_in, m_out, n_threads=8)
    parity.assert_bit_identical(a, b, "thread count")


def test_oracle_intersect_rays(oracle):
    g = load_golden("intersect_rays")
    parity.assert_bit_identical(oracle.intersect_rays(g["r1"], g["r2"]), g["pts"], "pairs")
    parity.assert_bit_identical(oracle.intersect_rays(g["axis_ray"], g["others"]), g["pts_axis"], "broadcast")
    parity.assert_bit_identical(oracle.intersect_rays(g["a"], g["b"]), g["pts_deg"], "degenerate")
    with pytest.raises(ValueError):
        oracle.intersect_rays(g["r1"][:3], g["r2"][:4])


def test_oracle_generators_match_reference(oracle):
    """NumPy restatement of get_ray_fan / get_collimated_rays; np.cos/np.sin are platform functions, so 2 ulp."""
    g = load_golden("generators")
    fan = oracle.source_rays("fan", 11, 7, 0.7, [0.1, 0.2, 0.3], (0, 0, 1), 0.6)
    col = oracle.source_rays("collimated", 9, 5, 4.0, [1, 2, 3], g["normal"], 0.5, b_start=0.25)
    coly = oracle.source_rays("collimated", 4, 3, 2.0, [0, 0, 0], (0, 1, 0), 0.5)
    for got, want in ((fan, g["fan"]), (col, g["collimated"]), (coly, g["collimated_y"])):
        np.testing.assert_allclose(got, want, rtol=0, atol=4e-16 * max(1.0, np.abs(want).max()))
    # slices of the index space agree with the full array
    part = oracle.source_rays("fan", 11, 7, 0.7, [0.1, 0.2, 0.3], (0, 0, 1), 0.6, first=13, count=20)
    assert np.array_equal(part, fan[13:33])


def test_oracle_sellmeier_matches_material_n(rtm, oracle):
    wl = np.linspace(0.35, 1.6, 4001)
    for name in ("Bk7", "Nlak22", "Nsf6ht", "Sf10", "FusedSilica", "Vacuum"):
        m = getattr(rtm, name)()
        got = oracle.sellmeier([m.b1, m.b2, m.b3], [m.c1, m.c2, m.c3], wl)
        parity.assert_bit_identical(got, m.n(wl), name)


def test_analytic_plano_convex_opl(rt, rtm, oracle):
    """known answer of scripts/2022_10_27_plano_convex_lens.py:39-44: OPL(h) through the singlet"""
    system, m_in, m_out, rays = systems.plano_convex(rt, rtm)
    hist = oracle.ray_trace(system, rays, m_in, m_out)
    k = 2 * np.pi / 0.5
    h = rays[:, 0]
    n, t0, t1, R, dz = 1.3, 2.679486355, 1, 100, 5
    sag = R - np.sqrt(R**2 - h**2)
    opl = dz + n * t0 + n * t1 - n * sag + sag / (np.sqrt(1 - n**2 * h**2 / R**2) * np.sqrt(R**2 - h**2) / R + n * h**2 / R**2)
    np.testing.assert_allclose(hist[-1, :, 6] / k, opl, rtol=0, atol=2e-14)


def test_analytic_perfect_lens_common_focus(rt, rtm, oracle):
    """scripts/2021_10_28_test_perfect_lens_phase.py: a tilted plane wave focuses to x = n1 f sin(theta), equal phases"""
    system, m_in, m_out, rays = systems.perfect_lens_phase(rt, rtm)
    hist = oracle.ray_trace(system, rays, m_in, m_out)
    x = hist[-1, :, 0]
    np.testing.assert_allclose(x, 1.1 * 4 * np.sin(10 * np.pi / 180), rtol=0, atol=1e-13)
    assert np.ptp(hist[-1, :, 6]) < 1e-9
