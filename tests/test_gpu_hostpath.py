"""
`-m gpu`: the host-buffer path (rtb_trace_host: the drop-in System.ray_trace call) at its edges --
  * N >= 2^25 rays, where the output pitch of the strided copy-out exceeds cudaDevAttrMaxPitch,
  * an error in the middle of the chunk pipeline: nothing may still be writing into the caller's array after the return,
  * batches whose wavelengths only differ in the sign of zero (grouping by bit pattern),
  * sweeps over more than RTB_MAX_WAVELENGTHS wavelengths (formula media: evaluated in the kernel; host-only media: groups),
  * user materials that answer an array with a scalar, wrong `out=` tensors.
"""
import numpy as np
import pytest

import parity
import systems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import ray_trace_pb_b200.device as dev
    return dev


def test_host_path_beyond_the_pitch_limit(rt, rtm, oracle):
    """2^25 + 1000 rays, keep='last' and a two-slab keep list (row pitch 2 GiB+: no strided 2-D copy possible)"""
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    n = (1 << 25) + 1000
    side = 4096
    base = systems.lattice_rays(side, 12.5, 0.0, 0.785, tilt=(0.002, -0.001))          # 2^24 rays
    rays = np.empty((n, 8))
    rays[: side * side] = base
    rays[side * side: 2 * side * side] = base[::-1]
    rays[2 * side * side:] = base[:1000]
    del base
    last = system.ray_trace(rays, vac, vac, keep="last")
    two = system.ray_trace(rays, vac, vac, keep=[12, 20])
    assert last.shape == (1, n, 8) and two.shape == (2, n, 8)
    parity.assert_bit_identical(two[1], last[0], "slab 20 of the keep list vs keep='last'")
    # spot checks against the oracle: the head, the seam between the chunks of the pipeline, the tail
    for lo in (0, (1 << 22) - 500, (1 << 24) - 500, (1 << 25) - 500, n - 1000):
        want = oracle.ray_trace(system, rays[lo:lo + 1000], vac, vac, n_threads=8)
        parity.assert_bit_identical(last[0, lo:lo + 1000], want[-1], f"rows {lo}..")
        parity.assert_bit_identical(two[0, lo:lo + 1000], want[12], f"slab 12 rows {lo}..")


def test_error_mid_pipeline_leaves_nothing_in_flight(rt, rtm):
    """rtb_tune('host_fail_chunk', 2): the call fails when it is about to launch its third chunk, after two chunks'
    copies have been enqueued; once it has returned, the caller's array is the caller's again"""
    import time
    from ray_trace_pb_b200 import _ffi, engine
    L = _ffi.lib()
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    mats = [vac] + list(system.materials) + [vac]
    rays = systems.lattice_rays(2048, 12.0, 0.0, 0.785)                     # 4.2M rays = 4+ chunks of 1M
    for pinned in (True, False):
        out = _ffi.pinned_empty((1, len(rays), 8)) if pinned else np.empty((1, len(rays), 8))
        out[:] = -1.0
        _ffi.check(L.rtb_tune(b"host_fail_chunk", 2))
        try:
            with pytest.raises(_ffi.RtbError, match="injected"):
                engine.trace_host(system.surfaces, mats, rays, keep="last", out=out)
        finally:
            _ffi.check(L.rtb_tune(b"host_fail_chunk", -1))
        out[:] = -7.0                    # ours again: a copy still in flight would overwrite some of this
        time.sleep(0.25)
        assert (out == -7.0).all(), f"writes landed after the error return (pinned={pinned})"
    # and the library is fine afterwards
    good = engine.trace_host(system.surfaces, mats, rays[:5000], keep="last")
    assert np.isfinite(good[0, :, 0]).sum() > 4000


def test_wavelengths_that_differ_in_the_sign_of_zero(rt, rtm, oracle):
    """nine bit patterns, eight values (+0.0 and -0.0), a host-only medium: the grouped route must terminate"""
    system, m_in, m_out, rays = systems.doublet_ebaf11(rt, rtm)
    wls = np.array([0.45, 0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.0, -0.0])
    big = np.tile(rays, (3, 1))
    big[:, 7] = wls[np.arange(len(big)) % len(wls)]
    got = system.ray_trace(big, m_in, m_out, keep="last")
    want = oracle.ray_trace(system, big, m_in, m_out, n_threads=4)[-1:]
    parity.assert_close_same_mask(got, want, rtol=1e-10, what="+-0 wavelengths")


def test_sweep_over_more_wavelengths_than_the_table_holds(rt, rtm, dev):
    import torch
    vac = rtm.Vacuum()
    wls = np.linspace(0.45, 0.85, 11)
    for builder in (systems.doublet_nlak22, systems.doublet_ebaf11):         # closed-formula media / a host-only medium
        system, m_in, m_out, _ = builder(rt, rtm)
        mats = [m_in] + list(system.materials) + [m_out]
        sources = [dev.RaySource.grid([0, 0, -10.0], 8.0, 41, float(w)) for w in wls]
        red = dev.Reducer(2 * len(system.surfaces) - 1, buckets=len(sources))
        out = dev.trace_sources(system.surfaces, mats, sources, keep="last", reducer=red)
        n = sources[0].n_rays
        for k, src in enumerate(sources):
            one = dev.Reducer(2 * len(system.surfaces) - 1)
            ref = dev.trace_source(system.surfaces, mats, src, keep="last", reducer=one)
            parity.assert_bit_identical(out[0, k * n:(k + 1) * n].cpu().numpy(), ref[0].cpu().numpy(),
                                        f"{builder.__name__} wavelength {k}")
            a, b = red.stats_t[k].cpu().numpy(), one.stats_t.cpu().numpy()
            assert a[0] == b[0]
            np.testing.assert_allclose(a[1:8], b[1:8], rtol=1e-10, atol=1e-9)
        torch.cuda.synchronize()


def test_scalar_answering_material_and_bad_out(rt, rtm, oracle, dev):
    import torch

    class Glassy(rtm.Material):                 # a user medium whose n() ignores the shape of its argument
        def __init__(self):
            pass

        def n(self, wavelength):
            return 1.62

    system, m_in, m_out, rays = systems.plano_convex(rt, rtm)
    system.materials[0] = Glassy()
    got = system.ray_trace(rays, m_in, m_out)
    system.materials[0] = rtm.Constant(1.62)
    want = oracle.ray_trace(system, rays, m_in, m_out)
    parity.assert_bit_identical(got, want, "scalar-answering material")

    mats = [m_in] + list(system.materials) + [m_out]
    src = dev.RaySource.grid([0, 0, -5.0], 5.0, 16, 0.5)
    for bad in (torch.empty((1, 255, 8), dtype=torch.float64, device="cuda"),
                torch.empty((1, 256, 8), dtype=torch.float32, device="cuda"),
                torch.empty((1, 256, 8), dtype=torch.float64)):
        with pytest.raises(ValueError):
            dev.trace_source(system.surfaces, mats, src, keep="last", out=bad)
        with pytest.raises(ValueError):
            dev.trace_sources(system.surfaces, mats, [src], keep="last", out=bad)


@pytest.mark.parametrize("n_rays", [1, 262144, 262145, 3 * 262144 + 17, 5 * 262144 - 1])
def test_lagged_pipeline_at_the_chunk_boundaries(n_rays, rt, rtm, oracle):
    """rtb_trace_host issues the copy-out of chunk c behind the copy-in of chunk c + 1 (four slots, 262 144-ray chunks
    for single-slab calls): one chunk, exactly one chunk, one ray more, a ragged tail, more chunks than slots --
    pageable and pinned buffers, keep='last' and a keep list, every row against the oracle"""
    from ray_trace_pb_b200 import _ffi, engine
    system = systems.relay10_system(rt, rtm)
    vac = rtm.Vacuum()
    materials = [vac] + list(system.materials) + [vac]
    rng = np.random.default_rng(n_rays)
    rays = np.zeros((n_rays, 8))
    rays[:, 0:2] = rng.uniform(-13.0, 13.0, (n_rays, 2))
    d = rng.standard_normal((n_rays, 3)) * np.array([0.02, 0.02, 0.0]) + np.array([0.0, 0.0, 1.0])
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 7] = 0.785
    want = oracle.ray_trace(system, rays, vac, vac, n_threads=8)
    # pageable in, result array from the library (pinned pool for big ones)
    last = system.ray_trace(rays, vac, vac, keep="last")
    parity.assert_bit_identical(last[0], want[-1], "pageable in, keep='last'")
    # pageable in AND out
    out = np.empty((1, n_rays, 8))
    engine.trace_host(system.surfaces, materials, rays, keep="last", out=out)
    parity.assert_bit_identical(out[0], want[-1], "pageable in and out")
    # pinned in and out, keep list (strided copy-out)
    pin_in = _ffi.pinned_empty(rays.shape)
    pin_in[:] = rays
    pin_out = _ffi.pinned_empty((3, n_rays, 8))
    pin_out[:] = -1.0
    engine.trace_host(system.surfaces, materials, pin_in, keep=[1, 12, 20], out=pin_out)
    parity.assert_bit_identical(pin_out, want[[1, 12, 20]], "pinned, keep list")
