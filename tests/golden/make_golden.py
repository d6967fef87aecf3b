"""
Generate the golden vectors for the hot path from the REFERENCE ITSELF.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py                      (everything)
    python tests/golden/make_golden.py --only-big NAME ...  (add / refresh single big-batch checksum pins)

What it writes (all committed):
  tests/golden/<case>.npz      rays_in (N,8), history (2S+1,N,8) from reference System.ray_trace, the system's
                               prescription as JSON, the unique wavelengths and the reference materials' n() on them
  tests/golden/checksums.json  sha256 of reference outputs for batches too large to store (platform-independent
                               lattice inputs, see tests/systems.py: lattice_rays), plus intersect_rays digests
  tests/golden/host_api.json   reference values of the host-only helpers (cardinal points, ABCD matrices, Seidel
                               sums, generators) used by tests/test_host_api.py

While generating, the CPU oracle (oracle/rt_oracle.c) is required to reproduce every reference output BIT FOR BIT;
the script aborts otherwise.  That is the pin the parity tests rest on.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def import_reference():
    """matplotlib is not installed in this image; the hot path never touches it."""
    for name in ("matplotlib", "matplotlib.figure", "matplotlib.axes", "matplotlib.axes._axes", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.figure"].Figure = object
    sys.modules["matplotlib.axes._axes"].Axes = object
    # make sure the *reference* package wins over the repo's drop-in shim of the same name
    for k in [k for k in sys.modules if k == "raytrace" or k.startswith("raytrace.")]:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference/src")
    import raytrace.raytrace as rt
    import raytrace.materials as rtm
    assert rt.__file__.startswith("/root/reference/"), rt.__file__
    sys.path.remove("/root/reference/src")
    return rt, rtm


def main(only_big=()):
    rt, rtm = import_reference()
    import systems
    import parity
    from oracle import oracle

    oracle.build(force=True)
    summary = {}

    # ------------------------------------------------------------------ per-case golden files
    for name, builder in ({} if only_big else systems.CASES).items():
        system, m_in, m_out, rays = builder(rt, rtm)
        with np.errstate(all="ignore"):
            hist = system.ray_trace(rays, m_in, m_out)
        # the oracle must agree with the reference bit for bit
        got = oracle.ray_trace(system, rays, m_in, m_out)
        parity.assert_bit_identical(got, hist, f"oracle vs reference, case {name}")

        materials = [m_in] + list(system.materials) + [m_out]
        uniq = np.unique(rays[:, 7])
        with np.errstate(all="ignore"):
            ntab = np.stack([np.asarray(m.n(uniq), dtype=float).reshape(-1) for m in materials], axis=1)
        desc = systems.describe_system(system, m_in, m_out)
        np.savez_compressed(HERE / f"{name}.npz", rays_in=rays, history=hist, system=json.dumps(desc),
                            unique_wavelengths=uniq, n_table=ntab)
        n_nan = int(np.isnan(hist[-1, :, 0]).sum())
        summary[name] = {"n_rays": int(rays.shape[0]), "n_surfaces": len(system.surfaces),
                         "nan_rays_at_end": n_nan, "sha256": parity.digest(hist)}
        print(f"{name:22s} N={rays.shape[0]:6d} S={len(system.surfaces):2d} NaN at end={n_nan:5d}  oracle == reference")

    # ------------------------------------------------------------------ input-shape variants of ray_trace
    system, m_in, m_out, rays = systems.plano_convex(rt, rtm, n_disps=9)
    single = system.ray_trace(rays[3], m_in, m_out)                       # (8,) input
    parity.assert_bit_identical(oracle.ray_trace(system, rays[3], m_in, m_out), single, "single ray")
    pre = np.stack((rays * 0.5, rays), axis=0)                            # (K,N,8) input: history is extended
    ext = system.ray_trace(pre, m_in, m_out)
    parity.assert_bit_identical(oracle.ray_trace(system, pre, m_in, m_out), ext, "history input")
    np.savez_compressed(HERE / "input_shapes.npz", rays=rays, single=single, pre=pre, ext=ext,
                        system=json.dumps(systems.describe_system(system, m_in, m_out)))

    # ------------------------------------------------------------------ big-batch checksums
    checks = {}
    big = {
        "relay10_lattice": (systems.relay10_system(rt, rtm), rtm.Vacuum(), rtm.Vacuum(),
                            systems.lattice_rays(320, 14.0, 0.0, 0.785, tilt=(0.004, -0.002))),
        "relay10_lattice_wide": (systems.relay10_system(rt, rtm), rtm.Vacuum(), rtm.Vacuum(),
                                 systems.lattice_rays(200, 27.0, 0.0, 0.785, tilt=(0.0, 0.0), converge=1e-4)),
        "plano_convex_lattice": (systems.plano_convex(rt, rtm)[0], rtm.Vacuum(), rtm.Vacuum(),
                                 systems.lattice_rays(256, 26.0, -5.0, 0.5)),
        "opm_lattice": (systems.opm_system(rt, rtm)[0], rtm.Constant(1.4), rtm.Vacuum(),
                        systems.lattice_rays(256, 1e-3, 1e-3, 532e-6, tilt=(0.0, 0.0), converge=600.0)),
        # 70 surfaces: longer than one kernel launch carries (the product chains two segments)
        "long_train_lattice": (systems.long_train_system(rt, rtm), rtm.Vacuum(), rtm.Vacuum(),
                               systems.lattice_rays(64, 15.5, 0.0, 0.5876, tilt=(0.002, -0.001))),
    }
    if only_big:
        # add / refresh single big-batch pins without rewriting the other (unchanged) golden files
        merged = json.loads((HERE / "checksums.json").read_text())
        big = {k: v for k, v in big.items() if k in only_big}
    for name, (system, m_in, m_out, rays) in big.items():
        with np.errstate(all="ignore"):
            hist = system.ray_trace(rays, m_in, m_out)
        got = oracle.ray_trace(system, rays, m_in, m_out, n_threads=8)
        parity.assert_bit_identical(got, hist, f"oracle vs reference, {name}")
        checks[name] = {"n_rays": int(rays.shape[0]), "sha256_history": parity.digest(hist),
                        "sha256_last": parity.digest(hist[-1]),
                        "nan_rays_at_end": int(np.isnan(hist[-1, :, 0]).sum()),
                        "system": systems.describe_system(system, m_in, m_out)}
        print(f"{name:22s} N={rays.shape[0]:6d} NaN at end={checks[name]['nan_rays_at_end']:6d}  oracle == reference")
    if only_big:
        merged["big"].update(checks)
        (HERE / "checksums.json").write_text(json.dumps(merged, indent=1))
        print("updated", sorted(checks), "in", HERE / "checksums.json")
        return

    # ------------------------------------------------------------------ random systems (checksums only)
    rand = {}
    for seed in range(systems.N_RANDOM_SYSTEMS):
        system, m_in, m_out, rays = systems.random_system(rt, rtm, seed)
        with np.errstate(all="ignore"):
            hist = system.ray_trace(rays, m_in, m_out)
        got = oracle.ray_trace(system, rays, m_in, m_out)
        parity.assert_bit_identical(got, hist, f"oracle vs reference, random system {seed}")
        kinds = "".join(type(s_).__name__[0] + type(s_).__name__[-1] for s_ in system.surfaces)
        rand[str(seed)] = {"sha256_history": parity.digest(hist), "n_surfaces": len(system.surfaces), "kinds": kinds,
                           "alive_at_end": int(np.isfinite(hist[-1, :, 0]).sum())}
    checks_random = rand
    print(f"random systems         {len(rand)} seeds, alive at end: "
          f"{[v['alive_at_end'] for v in rand.values()]}  oracle == reference")

    # ------------------------------------------------------------------ intersect_rays
    rng = np.random.default_rng(5)
    system, m_in, m_out, rays = systems.relay10_script(rt, rtm)
    hist = system.ray_trace(rays, m_in, m_out)
    r1 = hist[-2, 38:-1]
    r2 = hist[-2, 39:]
    with np.errstate(all="ignore"):
        pts = rt.intersect_rays(r1, r2)
        pts_axis = rt.intersect_rays(hist[-2, 9], hist[-2, 38:])
    parity.assert_bit_identical(oracle.intersect_rays(r1, r2), pts, "intersect_rays pairs")
    parity.assert_bit_identical(oracle.intersect_rays(hist[-2, 9], hist[-2, 38:]), pts_axis, "intersect_rays bcast")
    # degenerate configurations: parallel rays, rays in coordinate planes, skew rays
    a = np.zeros((64, 8))
    b = np.zeros((64, 8))
    a[:, 0:3] = rng.integers(-3, 4, (64, 3))
    b[:, 0:3] = rng.integers(-3, 4, (64, 3))
    a[:, 3:6] = rng.integers(-1, 2, (64, 3))
    b[:, 3:6] = rng.integers(-1, 2, (64, 3))
    a[:8, 3:6] = b[:8, 3:6]
    with np.errstate(all="ignore"):
        na = np.linalg.norm(a[:, 3:6], axis=1, keepdims=True)
        nb = np.linalg.norm(b[:, 3:6], axis=1, keepdims=True)
        a[:, 3:6] = np.where(na > 0, a[:, 3:6] / na, [0, 0, 1.0])
        b[:, 3:6] = np.where(nb > 0, b[:, 3:6] / nb, [0, 1.0, 0])
        pts_deg = rt.intersect_rays(a, b)
    parity.assert_bit_identical(oracle.intersect_rays(a, b), pts_deg, "intersect_rays degenerate")
    np.savez_compressed(HERE / "intersect_rays.npz", r1=r1, r2=r2, pts=pts, axis_ray=hist[-2, 9],
                        others=hist[-2, 38:], pts_axis=pts_axis, a=a, b=b, pts_deg=pts_deg)
    print("intersect_rays         oracle == reference")

    # ------------------------------------------------------------------ generators (NumPy restatement check)
    g1 = rt.get_ray_fan([0.1, 0.2, 0.3], 0.7, 11, 0.6, nphis=7, center_ray=(0, 0, 1))
    o1 = oracle.source_rays("fan", 11, 7, 0.7, [0.1, 0.2, 0.3], (0, 0, 1), 0.6)
    parity.assert_bit_identical(o1, g1, "fan generator")
    nrm = np.array([np.sin(0.2), 0, np.cos(0.2)])
    nrm = nrm / np.linalg.norm(nrm)
    g2 = rt.get_collimated_rays([1, 2, 3], 4.0, 9, 0.5, nphis=5, phi_start=0.25, normal=nrm)
    o2 = oracle.source_rays("collimated", 9, 5, 4.0, [1, 2, 3], nrm, 0.5, b_start=0.25)
    parity.assert_bit_identical(o2, g2, "collimated generator")
    g3 = rt.get_collimated_rays([0, 0, 0], 2.0, 4, 0.5, nphis=3, normal=(0, 1, 0))
    o3 = oracle.source_rays("collimated", 4, 3, 2.0, [0, 0, 0], (0, 1, 0), 0.5)
    parity.assert_bit_identical(o3, g3, "collimated generator, normal || y")
    np.savez_compressed(HERE / "generators.npz", fan=g1, collimated=g2, collimated_y=g3, normal=nrm)
    print("generators             oracle == reference")

    # ------------------------------------------------------------------ host-only helpers (reference values)
    host = {}
    kidger = rt.Doublet(rtm.Nsk11(), rtm.Nsf19(), radius_crown=64.1, radius_flint=-183.685,
                        radius_interface=-43.249, thickness_crown=3.5, thickness_flint=1.5,
                        aperture_radius=10.)
    kidger.set_aperture_stop(0)
    ab = kidger.seidel_third_order(0.5876, rtm.Vacuum(), rtm.Vacuum(), object_distance=np.inf,
                                   object_angle=0.01746)
    host["kidger_seidel"] = ab.tolist()
    ab2 = kidger.seidel_third_order(0.5876, rtm.Vacuum(), rtm.Vacuum(), object_distance=120.0, object_height=2.0)
    host["kidger_seidel_finite"] = ab2.tolist()
    host["kidger_abcd"] = kidger.get_ray_transfer_matrix(0.5876, rtm.Vacuum(), rtm.Vacuum()).tolist()
    cps = kidger.get_cardinal_points(0.5876, rtm.Constant(1.1), rtm.Constant(1.333))
    host["kidger_cardinal"] = [np.asarray(c, dtype=float).tolist() for c in cps]
    host["kidger_autofocus"] = {
        mode: np.asarray(kidger.auto_focus(0.5876, rtm.Vacuum(), rtm.Vacuum(), mode=mode), dtype=float).tolist()
        for mode in ("ray-fan", "collimated", "paraxial-focused", "paraxial-collimated")}
    host["kidger_gaussian_q"] = [[q.real, q.imag] for q in
                                 kidger.gaussian_paraxial(1j * 50.0, 0.5876, rtm.Vacuum(), rtm.Vacuum())]
    sysa = systems.achromat_imaging_system(rt, rtm)
    host["achromat_system"] = systems.describe_system(sysa, rtm.Vacuum(), rtm.Vacuum())
    host["achromat_aperture_stop"] = int(sysa.aperture_stop)
    host["achromat_surfaces_by_name"] = [int(v) for v in sysa.surfaces_by_name]
    host["achromat_names"] = list(sysa.names)
    host["achromat_seidel"] = sysa.seidel_third_order(0.635, rtm.Vacuum(), rtm.Vacuum(), object_height=5).tolist()
    host["materials_n"] = {}
    wls = np.array([0.4047, 0.4861, 0.5876, 0.6563, 0.785, 1.064])
    for cname in ("Vacuum", "FusedSilica", "Bk7", "Nbak4", "Nbaf10", "Nlak22", "Ebaf11", "Nsk11", "Sf10", "Nsf11",
                  "Nsf6", "Sf6", "Nsf6ht", "Sf2", "Nsf19"):
        m = getattr(rtm, cname)()
        host["materials_n"][cname] = {"n": np.asarray(m.n(wls), dtype=float).tolist(),
                                      "vd": None if m.vd is None else float(m.vd)}
    host["wavelengths"] = wls.tolist()
    angles, na = rt.ray_angle_about_axis(g1[:5], np.array([0, 0, 1.0]))
    host["ray_angle_about_axis"] = {"angles": angles.tolist(), "na": na.tolist()}
    dists, near = rt.dist_pt2plane(np.array([[1., 2, 3], [0, 0, -1]]), np.array([0, 0.6, 0.8]), np.array([0, 0, 1.]))
    host["dist_pt2plane"] = {"dists": dists.tolist(), "nearest": near.tolist()}
    p2p, ts = rt.propagate_ray2plane(g1[:6], np.array([0, 0.6, 0.8]), np.array([0, 0, 5.]), rtm.Bk7())
    np.savez_compressed(HERE / "ray2plane.npz", rays=g1[:6], out=p2p, ts=ts)
    (HERE / "host_api.json").write_text(json.dumps(host, indent=1))

    (HERE / "checksums.json").write_text(json.dumps({"cases": summary, "big": checks, "random": checks_random},
                                                    indent=1))
    print("wrote", HERE)


if __name__ == "__main__":
    # python tests/golden/make_golden.py [--only-big NAME ...]
    main(tuple(sys.argv[2:]) if len(sys.argv) > 2 and sys.argv[1] == "--only-big" else ())
