"""
GPU parity tests (`-m gpu`) of the lean trace kernel (csrc/trace_lean.cu: probe launch + warp-convergent in-place steps
+ whole-ray careful redo), which by default only takes launches of >= 32768 rays.  Here it is forced on for every
eligible launch (rtb_tune "lean_min_rays" = 0) and held against
  * the golden vectors written from the reference itself (final slab, bit for bit),
  * the sha256 pins of reference outputs of the 40 random systems and the 1e5-ray batches,
  * the CPU oracle on fuzzed bundles that exercise every route: lean steps, probe-selected zero-tolerant steps
    (sources on the first plane, beams along a flat's normal), meridional fans (exact-zero components ray after ray),
    NaN / inf launch rays, unlisted wavelengths, rays that die at every kind of surface,
  * the round-1 kernels (trace_f64.cu) on the same launches -- identical bits, identical reductions.
"""
import ctypes

import numpy as np
import pytest

import parity
import systems
from conftest import load_golden
from test_oracle_golden import BIG, _tables_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


@pytest.fixture(scope="module")
def dev():
    import ray_trace_pb_b200.device as dev
    return dev


@pytest.fixture(params=["probe_driven", "pure"])
def lean(request):
    """every eligible launch goes to the lean kernel for the duration of the test; `lean.off()` / `lean.on()` switch.
    Run twice: with the probe-driven kernels only (rtb_tune "lean_pure" = 0), and with the pure instantiations forced on
    for every system that has a lean step for each surface, whatever the bundle (= 2: no verdict, so bundles that the
    plain lean steps cannot take -- sources on a flat, beams along its normal -- go ray by ray through redo_ray)."""
    from ray_trace_pb_b200 import _ffi
    L = _ffi.lib()
    _ffi.check(L.rtb_tune(b"lean_pure", 2 if request.param == "pure" else 0))

    class Switch:
        @staticmethod
        def on():
            _ffi.check(L.rtb_tune(b"lean_min_rays", 0))

        @staticmethod
        def off():
            _ffi.check(L.rtb_tune(b"lean_min_rays", -1))

        @staticmethod
        def probe_counts(n_surfaces):
            buf = (ctypes.c_uint32 * (2 * n_surfaces))()
            _ffi.check(L.rtb_last_probe_counts(buf, n_surfaces))
            return np.array(buf[:], dtype=np.int64).reshape(n_surfaces, 2)

    Switch.on()
    _ffi.check(L.rtb_tune(b"keep_probe_counts", 1))
    yield Switch
    _ffi.check(L.rtb_tune(b"keep_probe_counts", 0))
    _ffi.check(L.rtb_tune(b"lean_min_rays", 32768))
    _ffi.check(L.rtb_tune(b"lean_pure", 1))


# ------------------------------------------------------------------------------------------------ reference pins
@pytest.mark.parametrize("name", sorted(systems.CASES))
def test_lean_golden_final_slab(name, rt, rtm, lean):
    g = load_golden(name)
    system, m_in, m_out = systems.rebuild_system(g["system"], rt, rtm)
    got = system.ray_trace(g["rays_in"], m_in, m_out, keep="last")
    want = g["history"][-1:]
    materials = [m_in] + system.materials + [m_out]
    if name in systems.POWER_DEPENDENT and not _tables_match(g, materials):
        parity.assert_close_same_mask(got, want, rtol=1e-10, what=name)
    else:
        parity.assert_bit_identical(got, want, name)


@pytest.mark.parametrize("name", sorted(BIG))
def test_lean_big_batch_checksum(name, rt, rtm, checksums, lean):
    ref = checksums["big"][name]
    system, m_in, m_out = systems.rebuild_system(ref["system"], rt, rtm)
    last = system.ray_trace(BIG[name](), m_in, m_out, keep="last")
    assert parity.digest(last[0]) == ref["sha256_last"]


@pytest.mark.parametrize("seed", range(systems.N_RANDOM_SYSTEMS))
def test_lean_random_systems(seed, rt, rtm, oracle, lean):
    """40 random systems (3-10 surfaces of all four kinds, decentred / tilted, random glasses): final slab == oracle"""
    system, m_in, m_out, rays = systems.random_system(rt, rtm, seed)
    got = system.ray_trace(rays, m_in, m_out, keep="last")
    want = oracle.ray_trace(system, rays, m_in, m_out, n_threads=8)[-1:]
    parity.assert_bit_identical(got, want, f"random system {seed}")


# ------------------------------------------------------------------------------------------------ bundles
def _fuzz(n, seed, spread=0.05, half=14.0, z0=-3.0, wavelengths=(0.785,)):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8))
    rays[:, 0:2] = rng.uniform(-half, half, (n, 2))
    rays[:, 2] = z0
    d = rng.standard_normal((n, 3)) * np.array([spread, spread, 0.0]) + np.array([0, 0, 1.0])
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 6] = rng.uniform(0, 50, n)
    rays[:, 7] = rng.choice(np.array(wavelengths), size=n)
    return rays


def _bundles():
    """name -> (system builder, rays): every route through the lean kernel"""
    out = {}
    # ordinary skew rays through the relay: every surface lean
    out["relay_skew"] = (systems.relay10_system, _fuzz(60_000, 1))
    # collimated along z: d x n has an exact-zero z component at the first lens, rounding-noise zeros in the coaxial part
    r = systems.lattice_rays(255, 13.0, 0.0, 0.785)
    out["relay_collimated"] = (systems.relay10_system, r)
    # meridional fan in the plane y = 0: exact zeros in every cross product, at every surface (probe -> general steps)
    r = _fuzz(40_000, 2)
    r[:, 1] = 0.0
    r[:, 4] = 0.0
    r[:, 3:6] /= np.linalg.norm(r[:, 3:6], axis=1, keepdims=True)
    out["relay_meridional"] = (systems.relay10_system, r)
    # NaN / inf launch rays, NaN and unlisted wavelengths sprinkled in (careful whole-ray redo)
    r = _fuzz(50_000, 3, wavelengths=(0.785, 0.532, 0.6328, 1.064, 0.405, 0.45, 0.5, 0.55, 0.6))
    rng = np.random.default_rng(4)
    r[rng.integers(0, len(r), 300)] = np.nan
    r[rng.integers(0, len(r), 300), 7] = np.nan
    r[rng.integers(0, len(r), 200), rng.integers(0, 7, 200)] = np.inf
    r[rng.integers(0, len(r), 200), rng.integers(0, 7, 200)] = np.nan
    out["relay_dirty"] = (systems.relay10_system, r)
    # wide bundle: rays die at apertures, miss spheres, go backwards
    out["relay_wide"] = (systems.relay10_system, _fuzz(60_000, 5, spread=0.4, half=40.0))
    return out


BUNDLES = _bundles()


@pytest.mark.parametrize("name", sorted(BUNDLES))
def test_lean_bundles_vs_oracle(name, rt, rtm, oracle, lean):
    builder, rays = BUNDLES[name]
    system = builder(rt, rtm)
    vac = rtm.Vacuum()
    got = system.ray_trace(rays, vac, vac, keep="last")
    want = oracle.ray_trace(system, rays, vac, vac, n_threads=8)[-1:]
    parity.assert_bit_identical(got, want, name)


def test_lean_config_systems_vs_oracle(rt, rtm, oracle, lean):
    """BASELINE configs 1, 2, 5 with the bundles their scripts launch: sources ON the first flat (t = +-0) and beams
    along its normal (d x n = 0) are the probe's business; the doublets behind are lean"""
    for builder in (systems.plano_convex, systems.doublet_nlak22, systems.achromat_imaging):
        system, m_in, m_out, rays = builder(rt, rtm)
        reps = max(1, 40_000 // len(rays))
        rng = np.random.default_rng(7)
        big = np.tile(rays, (reps, 1))
        big[len(rays):, 0:2] += rng.uniform(-1e-3, 1e-3, (len(big) - len(rays), 2))    # neighbours of the script's rays
        got = system.ray_trace(big, m_in, m_out, keep="last")
        want = oracle.ray_trace(system, big, m_in, m_out, n_threads=8)[-1:]
        parity.assert_bit_identical(got, want, builder.__name__)


# ------------------------------------------------------------------------------------------------ vs the round-1 kernels
def test_lean_equals_general_kernels_with_reductions(rt, rtm, oracle, dev, torch, lean):
    """final slab + a reduction at every slab the lean kernel takes (after a surface and at a surface): same final
    slab bits, same counts, same sums (1e-10: atomics reorder) as trace_f64.cu and as the oracle's history"""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    rays = BUNDLES["relay_wide"][1][:30_000].copy()
    rays[:20_000] = BUNDLES["relay_collimated"][1][:20_000]
    hist = oracle.ray_trace(system, rays, vac, vac, n_threads=8)
    d_rays = torch.from_numpy(rays).cuda()
    for slab in (1, 2, 5, 8, 11, 12, 19, 20):
        results = []
        for use_lean in (True, False):
            (lean.on if use_lean else lean.off)()
            red = dev.Reducer(slab, origin=(5.0, 0, 0), grid_n=64, half_width=15.0)
            out = dev.trace_tensor(system.surfaces, mats, d_rays, keep="last", wavelengths=[0.785], reducer=red)
            results.append((out.cpu().numpy(), red.stats_t.cpu().numpy(), red.grid.cpu().numpy()))
        lean.on()
        (o1, s1, g1), (o2, s2, g2) = results
        parity.assert_bit_identical(o1, o2, f"final slab, reduction at slab {slab}")
        parity.assert_bit_identical(o1, hist[-1:], f"final slab vs oracle, reduction at slab {slab}")
        want = oracle.reduce_stats(hist[slab], (5.0, 0, 0), (1, 0, 0), (0, 1, 0))
        assert s1[0] == want[0] == s2[0], f"count at slab {slab}"
        np.testing.assert_allclose(s1[1:8], want[1:8], rtol=1e-10, atol=1e-6)
        np.testing.assert_array_equal(s1[8:], want[8:])
        wg = oracle.reduce_grid(hist[slab], (5.0, 0, 0), (1, 0, 0), (0, 1, 0), 64, 15.0)
        np.testing.assert_array_equal(g1[2], wg[2])
        np.testing.assert_allclose(g1[:2], wg[:2], rtol=0, atol=1e-9 * max(1.0, np.abs(wg[2]).max()))


@pytest.fixture(params=["one_launch", "launch_per_source"])
def sweep_form(request):
    from ray_trace_pb_b200 import _ffi
    L = _ffi.lib()
    _ffi.check(L.rtb_tune(b"sweep_split_rays", 0 if request.param == "launch_per_source" else -1))
    yield request.param
    _ffi.check(L.rtb_tune(b"sweep_split_rays", 1 << 24))


def test_lean_sweep_equals_per_source_launches(rt, rtm, dev, torch, lean, sweep_form):
    """rtb_trace_sources (grid y = source, per-source probe counts and reduction buckets) through the lean kernel"""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    thetas = np.linspace(0, np.pi / 180, 5)
    sources = [dev.RaySource.grid([0, 0, 0], 12.0, 181, 0.785, normal=(np.sin(t), 0, np.cos(t))) for t in thetas]
    red = dev.Reducer(19, buckets=len(sources), grid_n=32, half_width=12.0)      # at the last surface
    out = dev.trace_sources(system.surfaces, mats, sources, keep="last", reducer=red)
    n = sources[0].n_rays
    for k, src in enumerate(sources):
        lean.off()
        one = dev.Reducer(19, grid_n=32, half_width=12.0)
        ref = dev.trace_source(system.surfaces, mats, src, keep="last", reducer=one)
        lean.on()
        parity.assert_bit_identical(out[0, k * n:(k + 1) * n].cpu().numpy(), ref[0].cpu().numpy(), f"source {k}")
        a, b = red.stats_t[k].cpu().numpy(), one.stats_t.cpu().numpy()
        assert a[0] == b[0]
        np.testing.assert_allclose(a[1:8], b[1:8], rtol=1e-10, atol=1e-6)
        np.testing.assert_array_equal(red.grid[k, 2].cpu().numpy(), one.grid[2].cpu().numpy())


# ------------------------------------------------------------------------------------------------ the probe
def test_probe_finds_the_surfaces_that_need_zero_forms(rt, rtm, dev, torch, lean):
    vac = rtm.Vacuum()
    # the relay under a collimated beam: nothing for the probe to find (the lean sphere step takes the exact-zero z
    # component of d x n in its stride)
    system = systems.relay10_system(rt, rtm)
    mats = [vac] + list(system.materials) + [vac]
    rays = torch.from_numpy(systems.lattice_rays(300, 12.0, 0.0, 0.785)).cuda()
    dev.trace_tensor(system.surfaces, mats, rays, keep="last", wavelengths=[0.785])
    counts = lean.probe_counts(len(system.surfaces))
    assert (counts[:, 0] > 1500).all() and (counts[:, 1] == 0).all(), counts
    # the plano-convex lens of config 1: the beam runs along the normal of both flats' common axis -- the first flat
    # sees d x n = 0 for every ray (the main launch then runs its zero-tolerant lean step there), the lens and the flat
    # behind it only for the few rays on the axis (one per azimuth) and in the planes x = 0 and y = 0
    system, m_in, m_out, _ = systems.plano_convex(rt, rtm)
    mats = [m_in] + list(system.materials) + [m_out]
    src = dev.RaySource.collimated([0, 0, -5], 20.0, 301, 0.5, nphis=301)
    dev.trace_source(system.surfaces, mats, src, keep="last")
    counts = lean.probe_counts(len(system.surfaces))
    assert counts[0, 1] == counts[0, 0] > 1500, counts
    assert (counts[1:, 1] * 50 <= counts[1:, 0]).all(), counts


# ------------------------------------------------------------------------------------------------ the verdict cache
@pytest.fixture()
def cache():
    """the default mode (rtb_tune "lean_pure" = 1) with the lean kernel taking every eligible launch"""
    from ray_trace_pb_b200 import _ffi
    L = _ffi.lib()
    _ffi.check(L.rtb_tune(b"lean_pure", 1))
    _ffi.check(L.rtb_tune(b"lean_min_rays", 0))
    yield L.rtb_pure_launch_count
    _ffi.check(L.rtb_tune(b"lean_min_rays", 32768))


def test_verdict_cache_picks_the_pure_kernel_from_the_second_launch(rt, rtm, oracle, dev, torch, cache):
    """launch 1 of a system: probe-driven kernel, its probe's counts are read back asynchronously; launches 2.. : the pure
    kernel.  Same bits from both, and from the oracle."""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    rays = _fuzz(50_000, 11, spread=0.03)
    surfaces = list(system.surfaces)        # (the `cache` fixture has emptied the cache)
    want = oracle.ray_trace(system, rays, vac, vac, n_threads=8)[-1:]
    d_rays = torch.from_numpy(rays).cuda()
    outs, pure = [], []
    for _ in range(4):
        before = cache()
        out = dev.trace_tensor(surfaces, mats, d_rays, keep="last", wavelengths=[0.785])
        torch.cuda.synchronize()
        pure.append(cache() - before)
        outs.append(out.cpu().numpy())
    assert pure == [0, 1, 1, 1], pure
    for o in outs:
        parity.assert_bit_identical(o, want, "relay, verdict cache")


def test_verdict_cache_zero_tolerant_flat_and_changing_bundles(rt, rtm, oracle, dev, torch, cache):
    """config 1's plano-convex lens: the beam runs along the first flat's normal, so the cached verdict puts the
    zero-tolerant lean flat into the pure kernel's runs; then the same system under bundles of another character (a
    skew fan, for which the verdict is stale: whatever fails is re-traced) -- always the oracle's bits; a key whose
    verdict keeps changing ends up with the probe-driven kernels."""
    system, m_in, m_out, _ = systems.plano_convex(rt, rtm)
    mats = [m_in] + list(system.materials) + [m_out]
    wl = 0.5
    along = systems.lattice_rays(200, 8.0, -5.0, wl)
    skew = _fuzz(40_000, 12, spread=0.02, half=8.0, z0=-5.0, wavelengths=(wl,))
    want = {"along": oracle.ray_trace(system, along, m_in, m_out, n_threads=8)[-1:],
            "skew": oracle.ray_trace(system, skew, m_in, m_out, n_threads=8)[-1:]}
    d = {"along": torch.from_numpy(along).cuda(), "skew": torch.from_numpy(skew).cuda()}

    def run(which):
        before = cache()
        out = dev.trace_tensor(system.surfaces, mats, d[which], keep="last", wavelengths=[wl])
        torch.cuda.synchronize()
        parity.assert_bit_identical(out.cpu().numpy(), want[which], f"plano-convex, {which} bundle")
        return cache() - before

    used = [run("along") for _ in range(3)]
    assert used == [0, 1, 1], used           # pure from the second launch on, with the zero-tolerant flat
    # alternate the bundles: every launch's verdict is the other bundle's; after three changes the key is left alone
    used = [run("skew" if k % 2 == 0 else "along") for k in range(10)]
    assert used[-3:] == [0, 0, 0], used


def test_verdict_cache_sources_and_sweeps(rt, rtm, dev, torch, cache):
    """on-device sources carry their description into the key: a sweep's verdict is its own"""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    # (fields off both coordinate planes: a beam tilted about y alone keeps its y = 0 row of rays in the meridional
    # plane, exact zeros at every surface -- 151 of 151^2 rays, above the probe's 1-in-200 threshold)
    thetas = np.linspace(0.001, np.pi / 180, 4)
    sources = [dev.RaySource.grid([0, 0, 0], 12.0, 151, 0.785, normal=(0.6 * np.sin(t), 0.8 * np.sin(t), np.cos(t)))
               for t in thetas]
    from ray_trace_pb_b200 import _ffi
    L = _ffi.lib()
    outs, used, kernels = [], [], []
    for _ in range(3):
        red = dev.Reducer(20, buckets=len(sources), grid_n=32, half_width=12.0)
        torch.cuda.synchronize()
        before, launched = cache(), L.rtb_launch_count()
        out = dev.trace_sources(system.surfaces, mats, sources, keep="last", reducer=red)
        torch.cuda.synchronize()
        used.append(cache() - before)
        kernels.append(L.rtb_launch_count() - launched)
        outs.append((out.cpu().numpy(), red.stats_t.cpu().numpy(), red.grid.cpu().numpy()))
    assert used == [0, 1, 1], used
    # the rays of an on-device source are a pure function of the key: once its verdict is in, the probe is not launched
    assert kernels == [2, 1, 1], kernels
    for o, s, g in outs[1:]:
        parity.assert_bit_identical(o, outs[0][0], "sweep, pure vs probe-driven")
        assert (s[:, 0] == outs[0][1][:, 0]).all()
        np.testing.assert_allclose(s[:, 1:8], outs[0][1][:, 1:8], rtol=1e-10, atol=1e-6)
        np.testing.assert_array_equal(g[:, 2], outs[0][2][:, 2])
