"""
Workload definitions shared by the golden-vector generator (run against the *reference* modules in the build
container) and by the parity tests (run against the product's host API).

Every builder takes the two API modules ``rt`` (raytrace.raytrace-like) and ``rtm`` (raytrace.materials-like) and
returns ``(system, initial_material, final_material, rays)``; the systems follow the reference's example scripts
(cited per builder, paths relative to /root/reference/scripts/).

``describe_system`` / ``rebuild_system`` move a system between the two APIs as plain numbers, so a golden file
pins the *prescription the reference actually traced*, independent of how well the product's paraxial helpers
(concatenate, get_cardinal_points) agree with the reference's in the last bit.
"""
from __future__ import annotations

import json

import numpy as np


# ----------------------------------------------------------------------------------------------------------
# neutral description <-> API objects
# ----------------------------------------------------------------------------------------------------------
def _surface_type(s):
    for klass in type(s).__mro__:
        if klass.__name__ in ("FlatSurface", "SphericalSurface", "PlaneMirror", "PerfectLens"):
            return klass.__name__
    raise TypeError(type(s))


def _vec(x):
    return [float(v) for v in np.asarray(x, dtype=float).reshape(-1)]


def describe_material(m):
    name = type(m).__name__
    if name == "Constant":
        return {"type": "Constant", "n": float(m._n)}
    if name == "Cauchy":
        return {"type": "Cauchy", "a": float(m.a), "b": float(m.b)}
    return {"type": name}


def describe_system(system, initial_material, final_material):
    surfaces = []
    for s in system.surfaces:
        t = _surface_type(s)
        d = {"type": t, "center": _vec(s.center), "paraxial_center": _vec(s.paraxial_center),
             "input_axis": _vec(s.input_axis), "output_axis": _vec(s.output_axis),
             "aperture_rad": float(s.aperture_rad)}
        if t == "SphericalSurface":
            d["radius"] = float(s.radius)
        else:
            d["normal"] = _vec(s.normal)
        if t == "PerfectLens":
            d["focal_len"] = float(s.focal_len)
            d["alpha"] = float(s.alpha)
        surfaces.append(d)
    return {"surfaces": surfaces,
            "materials": [describe_material(m) for m in system.materials],
            "initial_material": describe_material(initial_material),
            "final_material": describe_material(final_material)}


def make_cauchy(rtm):
    """A user-defined medium (n overridden), exercising the host refractive-index table."""

    class Cauchy(rtm.Material):
        def __init__(self, a, b):
            self.a = a
            self.b = b

        def n(self, wavelength):
            return self.a + self.b / np.asarray(wavelength) ** 2

    return Cauchy


def rebuild_material(d, rtm):
    if d["type"] == "Constant":
        return rtm.Constant(d["n"])
    if d["type"] == "Cauchy":
        return make_cauchy(rtm)(d["a"], d["b"])
    return getattr(rtm, d["type"])()


def rebuild_system(desc, rt, rtm):
    if isinstance(desc, np.ndarray):
        desc = desc.item()
    if isinstance(desc, (str, bytes, np.str_)):
        desc = json.loads(str(desc))
    surfaces = []
    for d in desc["surfaces"]:
        t = d["type"]
        if t == "FlatSurface":
            s = rt.FlatSurface(d["center"], d["normal"], d["aperture_rad"])
        elif t == "PlaneMirror":
            s = rt.PlaneMirror(d["center"], d["normal"], d["aperture_rad"])
        elif t == "SphericalSurface":
            s = rt.SphericalSurface(d["radius"], d["center"], d["aperture_rad"])
        elif t == "PerfectLens":
            s = rt.PerfectLens(d["focal_len"], d["center"], d["normal"], d["alpha"])
        else:
            raise TypeError(t)
        # reverse() and concatenate() edit these in place in the reference; carry the final values
        s.input_axis = np.array(d["input_axis"], dtype=float)
        s.output_axis = np.array(d["output_axis"], dtype=float)
        s.center = np.array(d["center"], dtype=float)
        s.paraxial_center = np.array(d["paraxial_center"], dtype=float)
        surfaces.append(s)
    system = rt.System(surfaces, [rebuild_material(m, rtm) for m in desc["materials"]])
    return system, rebuild_material(desc["initial_material"], rtm), rebuild_material(desc["final_material"], rtm)


# ----------------------------------------------------------------------------------------------------------
# BASELINE.json configs (small-N versions) and the reference's other example systems
# ----------------------------------------------------------------------------------------------------------
def plano_convex(rt, rtm, n_disps=101, nphis=1):
    """config 1: 2022_10_27_plano_convex_lens.py:14-34"""
    aperture_radius, t0, t1, rad_curv, n, dz = 25.4, 2.679486355, 1, 100, 1.3, 5
    system = rt.System([rt.FlatSurface([0, 0, 0], [0, 0, 1], aperture_radius),
                        rt.SphericalSurface.get_on_axis(-rad_curv, t0 + t1, aperture_radius),
                        rt.FlatSurface([0, 0, t0 + t1], [0, 0, 1], aperture_radius)],
                       [rtm.Constant(n), rtm.Vacuum()])
    rays = rt.get_collimated_rays([0, 0, -dz], aperture_radius, n_disps, 0.5, nphis=nphis)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def _doublet_with_focus(rt, rtm, doublet, wl_ref):
    """2022_08_04_ACT508-100-B.py:120-140 -- flat in front, doublet, flat at the paraxial focus"""
    cp = doublet.get_cardinal_points(wl_ref, rtm.Vacuum(), rtm.Vacuum())
    f2 = cp[1]
    system = rt.System([rt.FlatSurface([0, 0, -5.0], [0, 0, 1], 25.4)], [])
    system = system.concatenate(doublet, rtm.Vacuum())
    system = system.concatenate(rt.FlatSurface([0, 0, float(f2[2])], [0, 0, 1], 25.4), rtm.Vacuum())
    return system


def doublet_ebaf11(rt, rtm, n_disps=41, nphis=6):
    """config 2 (script's active lens): AC508-075-A-ML, Ebaf11/Nsf11, 3 wavelengths in ONE batch"""
    doublet = rt.Doublet(rtm.Ebaf11(), rtm.Nsf11(), radius_crown=50.8, radius_flint=-247.7,
                         radius_interface=-41.7, thickness_crown=20., thickness_flint=3.,
                         aperture_radius=25.4, input_collimated=True, names="AC508-075-A-ML")
    system = _doublet_with_focus(rt, rtm, doublet, 0.5876)
    rays = np.concatenate([rt.get_collimated_rays([0, 0, -10], 22.0, n_disps, wl, nphis=nphis)
                           for wl in (0.4861, 0.5876, 0.6563)], axis=0)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def doublet_nlak22(rt, rtm, n_disps=41, nphis=6):
    """config 2 (BASELINE-named lens): AC508-100-B, Nlak22/Nsf6ht, 2022_08_04_ACT508-100-B.py:58-72"""
    doublet = rt.Doublet(rtm.Nlak22(), rtm.Nsf6ht(), radius_crown=65.8, radius_flint=-280.6,
                         radius_interface=-56, thickness_crown=13.0, thickness_flint=2.0,
                         aperture_radius=25.4, input_collimated=True, names="AC508-100-B")
    system = _doublet_with_focus(rt, rtm, doublet, 0.855)
    rays = np.concatenate([rt.get_collimated_rays([0, 0, -10], 24.0, n_disps, wl, nphis=nphis)
                           for wl in (0.7065, 0.855, 1.015)], axis=0)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def relay10_system(rt, rtm, offset=5.0):
    """config 3 / config 5: the 10-surface relay of 2022_08_24_relay_astigmatism.py:13-80"""
    t100c, r100c, r100i, t100f, r100f, bfl100 = 13.0, 65.8, -56., 2.0, -280.6, 91.5
    t180c, r180c, r180i, t180f, r180f, bfl180 = 9.5, 144.4, -115.4, 4.0, -328.2, 173.52
    t300c, r300c, r300i, t300f, bfl300, efl300 = 9.0, 167.7, -285.8, 4.0, 289.81, 300
    radius = 25.4
    z180 = 10
    z100 = (t180c + t180f) + bfl180 + 101.5
    z300 = z100 + (t100c + t100f) + bfl100 + efl300
    zend = z300 + (t300c + t300f) + bfl300
    S = rt.SphericalSurface
    system = rt.System(
        [S(r180c, [offset, 0, z180 + np.abs(r180c)], radius),
         S(r180i, [offset, 0, z180 + t180c - np.abs(r180i)], radius),
         S(r180f, [offset, 0, z180 + t180c + t180f - np.abs(r180f)], radius),
         S(-r100f, [offset, 0, z100 + np.abs(r100f)], radius),
         S(-r100i, [offset, 0, z100 + t100f + np.abs(r100i)], radius),
         S(-r100c, [offset, 0, z100 + t100f + t100c - np.abs(r100c)], radius),
         S.get_on_axis(r300c, z300, radius),
         S.get_on_axis(r300i, z300 + t300c, radius),
         rt.FlatSurface([0, 0, z300 + t300c + t300f], [0, 0, 1], radius),
         rt.FlatSurface([0, 0, zend], [0, 0, 1], radius)],
        [rtm.Nlak22(), rtm.Nsf6(), rtm.Constant(1),
         rtm.Nsf6ht(), rtm.Nlak22(), rtm.Constant(1),
         rtm.Nlak22(), rtm.Nsf6(), rtm.Constant(1)])
    return system


def long_train_system(rt, rtm):
    """
    70 surfaces -- more than one kernel launch carries (RTB_MAX_SURFACES = 64), so the product traces it in two chained
    segments; the reference's loop has no such limit.  33 weak singlets of alternating sign (66 spherical surfaces,
    three glasses), a window whose back face is tilted, a perfect lens and a screen in its back focal plane.
    """
    S = rt.SphericalSurface
    glasses = [rtm.Bk7(), rtm.Sf10(), rtm.FusedSilica()]
    surfaces, mats = [], []
    z = 10.0
    for k in range(33):
        sign = 1.0 if k % 2 == 0 else -1.0
        R = 400.0 + 10.0 * k
        surfaces.append(S.get_on_axis(sign * R, z, 15.0))
        mats.append(glasses[k % 3])
        surfaces.append(S.get_on_axis(-sign * R, z + 3.0, 15.0))
        mats.append(rtm.Constant(1) if k % 4 else rtm.Vacuum())
        z += 8.0
    surfaces.append(rt.FlatSurface([0, 0, z + 5.0], [0, 0, 1], 15.0))
    mats.append(rtm.Bk7())
    surfaces.append(rt.FlatSurface([0, 0, z + 8.0], [np.sin(0.05), 0, np.cos(0.05)], 15.0))
    mats.append(rtm.Vacuum())
    surfaces.append(rt.PerfectLens(50.0, [0, 0, z + 30.0], [0, 0, 1], 0.25))
    mats.append(rtm.Vacuum())
    surfaces.append(rt.FlatSurface([0, 0, z + 80.0], [0, 0, 1], 15.0))
    return rt.System(surfaces, mats)


def relay10(rt, rtm, n_disps=15, nphis=12, n_fields=3, beam_rad=12.0):
    """config 3: tilted collimated bundles through the relay (field angles 0..1 degree)"""
    system = relay10_system(rt, rtm)
    bundles = []
    for theta in np.linspace(0, 1 * np.pi / 180, n_fields):
        normal = np.array([np.sin(theta), 0, np.cos(theta)])
        normal = normal / np.linalg.norm(normal)
        bundles.append(rt.get_collimated_rays([0, 0, 0], beam_rad, n_disps, 0.785, nphis=nphis, normal=normal))
    return system, rtm.Vacuum(), rtm.Vacuum(), np.concatenate(bundles, axis=0)


def relay10_script(rt, rtm):
    """the exact ray set of 2022_08_24_relay_astigmatism.py:82-85 (19 + 19 + 1900 rays)"""
    wavelength, nrays = 0.785, 19
    beam_rad = 20e-3 * np.sqrt(1 + (3 / (np.pi * 20e-3**2 / (wavelength * 1e-3)))**2)
    system = relay10_system(rt, rtm)
    rays = np.concatenate((rt.get_collimated_rays([0, 0, 0], beam_rad, nrays, wavelength),
                           rt.get_collimated_rays([0, 0, 0], beam_rad, nrays, wavelength, phi_start=np.pi / 2),
                           rt.get_collimated_rays([0, 0, 0], beam_rad, nrays, wavelength, nphis=100)), axis=0)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def opm_system(rt, rtm):
    """config 4: 11-surface ideal OPM, 2022_01_25_ray_trace_ideal_opm.py:10-80"""
    aperture_rad = 2
    n1, na1, mag1 = 1.4, 1.35, 100
    alpha1 = np.arcsin(na1 / n1)
    f1 = 200 / mag1
    n2, na2, mag2 = 1, 0.95, 40
    alpha2 = np.arcsin(na2 / n2)
    f2 = 200 / mag2
    r2 = na2 * f2
    theta = 30 * np.pi / 180
    n3, na3 = 1.51, 1
    alpha3 = np.arcsin(na3 / n3)
    f3 = 200 / 100
    o3_normal = np.array([-np.sin(theta), 0, np.cos(theta)])
    ft1 = 200
    ft2 = ft1 / f1 * f2 / n1
    ft3 = 200
    p_o1 = n1 * f1
    p_pupil_o1 = p_o1 + f1
    p_t1 = p_o1 + f1 + ft1
    p_t2 = p_t1 + ft1 + ft2
    p_pupil_o2 = p_t2 + ft2
    p_o2 = p_t2 + ft2 + f2
    p_remote_focus = p_o2 + n2 * f2
    p_o3 = np.array([0, 0, p_remote_focus]) + n3 * f3 * o3_normal
    p_pupil_o3 = p_o3 + f3 * o3_normal
    p_t3 = p_o3 + (f3 + ft3) * o3_normal
    p_imag = p_t3 + ft3 * o3_normal
    system = rt.System([rt.PerfectLens(f1, [0, 0, p_o1], [0, 0, 1], alpha1),
                        rt.FlatSurface([0, 0, p_pupil_o1], [0, 0, 1], n1 * f1),
                        rt.PerfectLens(ft1, [0, 0, p_t1], [0, 0, 1], alpha1),
                        rt.PerfectLens(ft2, [0, 0, p_t2], [0, 0, 1], alpha2),
                        rt.FlatSurface([0, 0, p_pupil_o2], [0, 0, 1], n2 * f2),
                        rt.PerfectLens(f2, [0, 0, p_o2], [0, 0, 1], alpha2),
                        rt.FlatSurface([0, 0, p_remote_focus], o3_normal, r2),
                        rt.PerfectLens(f3, p_o3, o3_normal, alpha3),
                        rt.FlatSurface(p_pupil_o3, o3_normal, f3 * n3),
                        rt.PerfectLens(ft3, p_t3, o3_normal, alpha3),
                        rt.FlatSurface(p_imag, o3_normal, aperture_rad)],
                       [rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum(),
                        rtm.Constant(n2), rtm.Constant(n3), rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum()])
    return system, rtm.Constant(n1), rtm.Vacuum(), alpha1, theta


def opm(rt, rtm, n_thetas=41, nphis=12):
    system, m_in, m_out, alpha1, theta = opm_system(rt, rtm)
    dx = 0.001
    rays = rt.get_ray_fan([dx, dx, dx * np.tan(theta)], alpha1, n_thetas, 532e-6, nphis=nphis)
    return system, m_in, m_out, rays


def achromat_imaging_system(rt, rtm, wlen=0.635):
    """config 5: 9-surface achromat 4f imaging system, 2024_08_08_achromat_imaging.py:11-70"""
    kw = dict(radius_crown=50.8, radius_flint=-247.7, radius_interface=-41.7, thickness_crown=20.,
              thickness_flint=3., aperture_radius=25.4, names="AC508-075-A-ML")
    l1 = rt.Doublet(rtm.Ebaf11(), rtm.Nsf11(), input_collimated=False, **kw)
    l2 = rt.Doublet(rtm.Ebaf11(), rtm.Nsf11(), input_collimated=True, **kw)
    vac = rtm.Vacuum
    cp1 = l1.get_cardinal_points(wlen, vac(), vac())
    f1_left = cp1[0][-1]
    f1_right = cp1[1][-1]
    wd_right = f1_right - l1.surfaces[-1].paraxial_center[-1]
    system = rt.System([rt.FlatSurface([0, 0, 0], [0, 0, 1], 25.4)], [])
    system = system.concatenate(l1, vac(), -f1_left)
    d = l2.find_paraxial_collimated_distance(l2, wlen, vac(), vac(), vac())
    system = system.concatenate(rt.FlatSurface([0, 0, 0], [0, 0, 1], 25.4), vac(), wd_right)
    ind_pupil = len(system.surfaces) - 1
    system = system.concatenate(l2, vac(), d - wd_right)
    c2 = l2.get_cardinal_points(wlen, vac(), vac())
    wd2 = c2[1][2] - l2.surfaces[-1].paraxial_center[2]
    system = system.concatenate(rt.FlatSurface([0, 0, 0], [0, 0, 1], 25.4), vac(), wd2)
    system.set_aperture_stop(ind_pupil)
    return system


def achromat_imaging(rt, rtm, wlen=0.635, n_heights=9, nrays=7, nphis=4):
    system = achromat_imaging_system(rt, rtm, wlen)
    fans = [rt.get_ray_fan(np.array([h, 0, 0]), 4 * np.pi / 180, nrays, wlen, nphis=nphis)
            for h in np.linspace(0, 16, n_heights)]
    return system, rtm.Vacuum(), rtm.Vacuum(), np.concatenate(fans, axis=0)


def mirrors(rt, rtm, n_thetas=9, nphis=5):
    """2021_07_25_mirror.py:9-17, with a wider fan so some rays miss the second mirror"""
    theta = np.pi / 4 - np.pi / 30
    system = rt.System([rt.PlaneMirror([0, 0, 30], [-np.sin(theta), 0, -np.cos(theta)], 25),
                        rt.PlaneMirror([-50, 0, 30], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)], 25),
                        rt.FlatSurface([-50, 0, 60], [0, 0, 1], 25)],
                       [rtm.Vacuum(), rtm.Vacuum()])
    rays = rt.get_ray_fan([0, 0, 0], 25 * np.pi / 180, n_thetas, 0.785, nphis=nphis)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def perfect_lens_phase(rt, rtm, nrays=7, nphis=1):
    """2021_10_28_test_perfect_lens_phase.py:12-41 -- known answer: common focus, identical phase"""
    wavelength, aperture, n1, n2, f = 0.785, 10, 1.1, 1.3, 4
    alpha = np.arcsin(1 / n1)
    system = rt.System([rt.FlatSurface([0, 0, 0], [0, 0, 1], aperture),
                        rt.PerfectLens(f, [0, 0, n1 * f], [0, 0, 1], alpha),
                        rt.FlatSurface([0, 0, n1 * f + n2 * f], [0, 0, 1], aperture)],
                       [rtm.Constant(n1), rtm.Constant(n2)])
    angle = 10 * np.pi / 180
    rays = rt.get_collimated_rays([0, 0, -1], 3, nrays, wavelength, nphis=nphis,
                                  normal=[np.sin(angle), 0, np.cos(angle)])
    return system, rtm.Constant(n1), rtm.Constant(n2), rays


def perfect_imaging(rt, rtm, n_thetas=31, nphis=9, dz=0.002):
    """2022_02_06_perfect_imaging_system_psf.py: two perfect lenses + pupil and image flats, defocused source"""
    wavelength = 0.532e-3
    n1, na, f1, f2 = 1.0, 0.3, 3.0, 30.0
    alpha = np.arcsin(na / n1)
    system = rt.System([rt.PerfectLens(f1, [0, 0, f1], [0, 0, 1], alpha),
                        rt.FlatSurface([0, 0, 2 * f1], [0, 0, 1], 3 * f1),
                        rt.PerfectLens(f2, [0, 0, 2 * f1 + f2], [0, 0, 1], alpha),
                        rt.FlatSurface([0, 0, 2 * f1 + 2 * f2], [0, 0, 1], 10.)],
                       [rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum()])
    rays = rt.get_ray_fan([0.001, -0.0005, dz], 1.2 * alpha, n_thetas, wavelength, nphis=nphis)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def perfect_imaging_in_focus(rt, rtm, n_thetas=21, nphis=8):
    """the textbook bundle: point source exactly at the front focal point (focal-plane height exactly 0), hence a beam
    exactly along the axis into the second lens (transverse direction exactly 0)"""
    wavelength = 0.532e-3
    na, f1, f2 = 0.3, 3.0, 30.0
    alpha = np.arcsin(na)
    system = rt.System([rt.PerfectLens(f1, [0, 0, f1], [0, 0, 1], alpha),
                        rt.FlatSurface([0, 0, 2 * f1], [0, 0, 1], 3 * f1),
                        rt.PerfectLens(f2, [0, 0, 2 * f1 + f2], [0, 0, 1], alpha),
                        rt.FlatSurface([0, 0, 2 * f1 + 2 * f2], [0, 0, 1], 10.)],
                       [rtm.Vacuum(), rtm.Vacuum(), rtm.Vacuum()])
    rays = rt.get_ray_fan([0.0, 0.0, 0.0], 1.2 * alpha, n_thetas, wavelength, nphis=nphis)
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def retro_mirror(rt, rtm, n=15, nphis=8):
    """a collimated beam sent straight back by a mirror (exactly normal incidence on mirror and windows), then a
    second, tilted bundle"""
    system = rt.System([rt.FlatSurface([0, 0, 0], [0, 0, 1], 25.0),
                        rt.PlaneMirror([0, 0, 30], [0, 0, -1], 25.0),
                        rt.FlatSurface([0, 0, 10], [0, 0, -1], 25.0),
                        rt.SphericalSurface(-80.0, [0, 0, 5.0 - 80.0], 25.0, input_axis=(0, 0, -1))],
                       [rtm.Bk7(), rtm.Bk7(), rtm.Constant(1.3)])
    a = rt.get_collimated_rays([0, 0, -5], 12.0, n, 0.6, nphis=nphis)
    ang = 0.05
    b = rt.get_collimated_rays([0, 0, -5], 12.0, n, 0.6, nphis=nphis, normal=[np.sin(ang), 0, np.cos(ang)])
    return system, rtm.Vacuum(), rtm.Vacuum(), np.concatenate((a, b), axis=0)


def cauchy_singlet(rt, rtm, n=200, seed=7):
    """a user-defined medium (overridden n) with one wavelength per ray drawn from 5 lines"""
    Cauchy = make_cauchy(rtm)
    system = rt.System([rt.SphericalSurface.get_on_axis(60.0, 0.0, 20.0),
                        rt.SphericalSurface.get_on_axis(-60.0, 8.0, 20.0),
                        rt.FlatSurface([0, 0, 60.0], [0, 0, 1], 30.0)],
                       [Cauchy(1.5046, 0.0042), rtm.Vacuum()])
    rng = np.random.default_rng(seed)
    rays = rt.get_collimated_rays([0, 0, -5], 19.0, n, 0.5, nphis=1, phi_start=0.3)
    rays[:, 7] = rng.choice(np.array([0.45, 0.5, 0.55, 0.6, 0.65]), size=n)
    return system, rtm.Vacuum(), rtm.Bk7(), rays


def edge_mix(rt, rtm, n=600, seed=11):
    """
    Random rays thrown at a short mixed system so that every invalidation path fires: sphere misses, both roots
    negative, back-propagation at a flat, not-incoming cull, TIR (glass -> vacuum at steep angles), aperture
    culls, rays parallel to a plane, NaN rows and NaN wavelengths in the input, infinite aperture.
    """
    system = rt.System([rt.FlatSurface([0, 0, 0], [0, 0, 1], np.inf),
                        rt.SphericalSurface.get_on_axis(30.0, 5.0, 12.0),
                        rt.SphericalSurface.get_on_axis(-25.0, 14.0, 12.0),
                        rt.FlatSurface([0, 0, 20.0], [0.1, 0, np.sqrt(1 - 0.01)], 15.0),
                        rt.PlaneMirror([0, 0, 40.0], [0, np.sin(0.4), -np.cos(0.4)], 18.0),
                        rt.FlatSurface([0, 30.0, 25.0], [0, 1, 0], 40.0)],
                       [rtm.Constant(1.0), rtm.Sf10(), rtm.Constant(1.7), rtm.Vacuum(), rtm.Bk7()])
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8))
    rays[:, 0:2] = rng.uniform(-14, 14, (n, 2))
    rays[:, 2] = rng.uniform(-8, -1, n)
    d = rng.standard_normal((n, 3)) * np.array([0.45, 0.45, 0.2]) + np.array([0, 0, 1.0])
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3:6] = d
    rays[:, 6] = rng.uniform(0, 10, n)
    rays[:, 7] = rng.choice(np.array([0.405, 0.532, 0.785, 1.064]), size=n)
    # hand-made special rows
    rays[0, 3:6] = [0, 0, 1]                      # on-axis, normal incidence everywhere
    rays[0, 0:2] = 0
    rays[1, 3:6] = [0, 0, -1]                     # travelling backwards
    rays[2, 3:6] = [1, 0, 0]                      # parallel to the first plane
    rays[3, :] = np.nan                           # dead on arrival
    rays[4, 7] = np.nan                           # valid geometry, NaN wavelength
    rays[5, 2] = 3.0                              # starts behind the first flat
    rays[6, 0:3] = [0, 0, 0]                      # starts exactly on the first flat
    rays[6, 3:6] = [0, 0, 1]
    rays[7, 3:6] = [0.6, 0, 0.8]                  # steep: TIR candidates
    rays[8, 3:6] = [0, -0.8, 0.6]
    rays[9, 6] = np.nan                           # NaN phase only
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def invalid_inputs(rt, rtm, n=1200, seed=21):
    """launch rays with inf / NaN in single columns, through the edge-mix system (users can pass these; the reference
    itself only ever blanks x, y, z together)"""
    system, m_in, m_out, _ = edge_mix(rt, rtm)
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8))
    rays[:, 0:2] = rng.uniform(-4, 4, (n, 2))
    rays[:, 2] = rng.uniform(-6, 0, n)
    d = rng.standard_normal((n, 3)) * np.array([0.2, 0.2, 0.1]) + np.array([0, 0, 1.0])
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 6] = rng.uniform(0, 100, n)
    rays[:, 7] = rng.choice(np.array([0.405, 0.532, 0.785, 1.064]), size=n)
    for col in range(8):
        rows = rng.integers(0, n, 45)
        rays[rows[:25], col] = np.nan
        rays[rows[25:35], col] = np.inf
        rays[rows[35:], col] = -np.inf
    return system, m_in, m_out, rays


def reversed_doublet(rt, rtm, n_disps=21, nphis=4):
    """System.reverse() flips input_axis only (raytrace.py:402-415): trace right-to-left through a doublet"""
    doublet = rt.Doublet(rtm.Nbak4(), rtm.Sf10(), radius_crown=61.5, radius_flint=-128.2, radius_interface=-44.6,
                         thickness_crown=8.0, thickness_flint=2.5, aperture_radius=12.7, input_collimated=True)
    system = doublet.reverse()
    rays = rt.get_collimated_rays([0, 0, 40.0], 11.0, n_disps, 0.633, nphis=nphis, normal=[0, 0, -1])
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


def random_system(rt, rtm, seed: int, n_rays: int = 1500):
    """
    A random sequential system (3-10 surfaces of all four kinds, random glasses, decentred / tilted elements) and a
    random ray bundle.  Built from a seeded PCG64 stream and + - * / sqrt only, so it is the same on every platform:
    make_golden.py records the sha256 of the REFERENCE's history for each seed, tests re-derive it.
    """
    rng = np.random.default_rng(1000 + seed)
    glasses = [rtm.Vacuum, lambda: rtm.Constant(1.33), rtm.Bk7, rtm.Sf10, rtm.Nlak22, rtm.FusedSilica, rtm.Nsf6,
               lambda: rtm.Constant(1.0)]

    def unit(v):
        v = np.asarray(v, dtype=float)
        return v / np.sqrt((v * v).sum())

    n_surf = int(rng.integers(3, 11))
    surfaces, z = [], 0.0
    for k in range(n_surf):
        kind = rng.choice(["sphere", "sphere", "sphere", "flat", "flat", "lens", "mirror"]) if k == n_surf - 1 else \
            rng.choice(["sphere", "sphere", "sphere", "flat", "flat", "lens"])
        off = rng.uniform(-1.5, 1.5, 2) * (rng.random() < 0.5)
        aperture = float(rng.uniform(5.0, 15.0))
        if kind == "sphere":
            radius = float(rng.uniform(20.0, 200.0) * (1 if rng.random() < 0.5 else -1))
            surfaces.append(rt.SphericalSurface(radius, [off[0], off[1], z + radius], aperture))
        else:
            tilt = rng.uniform(-0.08, 0.08, 2) * (rng.random() < 0.5)
            normal = unit([tilt[0], tilt[1], 1.0]) if tilt.any() else np.array([0.0, 0.0, 1.0])
            if kind == "flat":
                surfaces.append(rt.FlatSurface([off[0], off[1], z], normal, aperture))
            elif kind == "mirror":
                surfaces.append(rt.PlaneMirror([off[0], off[1], z], normal, aperture))
            else:
                surfaces.append(rt.PerfectLens(float(rng.uniform(20.0, 100.0)), [off[0], off[1], z], normal,
                                               float(rng.uniform(0.2, 0.6))))
        z += float(rng.uniform(2.0, 30.0))
    pick = lambda: glasses[int(rng.integers(0, len(glasses)))]()
    system = rt.System(surfaces, [pick() for _ in range(n_surf - 1)])
    m_in, m_out = pick(), pick()

    rays = np.zeros((n_rays, 8))
    r = 8.0 * np.sqrt(rng.random(n_rays))
    c = unit(rng.standard_normal((2, n_rays)).T[0])  # unused draw keeps the stream simple to reason about
    del c
    ang = rng.standard_normal((n_rays, 2))
    ang = ang / np.sqrt((ang * ang).sum(axis=1, keepdims=True))
    rays[:, 0] = r * ang[:, 0]
    rays[:, 1] = r * ang[:, 1]
    rays[:, 2] = -10.0
    d = rng.standard_normal((n_rays, 3)) * np.array([0.06, 0.06, 0.0]) + np.array([0.0, 0.0, 1.0])
    rays[:, 3:6] = d / np.sqrt((d * d).sum(axis=1, keepdims=True))
    rays[:, 6] = rng.uniform(0, 50, n_rays)
    rays[:, 7] = rng.choice(rng.uniform(0.4, 1.1, 3), size=n_rays)
    rays[rng.integers(0, n_rays, 5)] = np.nan
    return system, m_in, m_out, rays


def launched_on_plane(rt, rtm, which, n=400, seed=77):
    """
    Rays that start exactly on the first (plane) surface, t = +-0 there, as when a script puts its source on the flat
    at z = 0; zero components of both signs in positions, directions, phases and normals, so the sign of every exact
    zero is exercised.  `which`: 0 flat +z, 1 flat with normal (-0, 0, -1), 2 tilted flat, 3 mirror, 4 perfect lens.
    """
    rng = np.random.default_rng(seed + which)
    zero = np.array([0.0, -0.0])
    tilt = np.array([np.sin(0.2), 0.0, np.cos(0.2)])
    first = [lambda: rt.FlatSurface([0, 0, 0], [0, 0, 1], 30.0),
             lambda: rt.FlatSurface([0, 0, 0], [-0.0, 0.0, -1.0], 30.0),
             lambda: rt.FlatSurface([0, 0, 0], tilt, 30.0),
             lambda: rt.PlaneMirror([0, 0, 0], [0.0, -0.0, 1.0], 30.0),
             lambda: rt.PerfectLens(20.0, [0, 0, 0], [0, 0, 1], 0.6)][which]()
    back = which in (1, 3)   # light leaves these towards -z
    system = rt.System([first, rt.SphericalSurface.get_on_axis(40.0, -5.0 if back else 5.0, 25.0),
                        rt.FlatSurface([0, 0, -9.0 if back else 9.0], [0, 0, -1.0 if back else 1.0], 25.0)],
                       [rtm.Bk7(), rtm.Constant(1.5)])
    if back:
        system.surfaces[1].input_axis = -system.surfaces[1].input_axis
        system.surfaces[1].output_axis = -system.surfaces[1].output_axis
    rays = np.zeros((n, 8))
    rays[:, 0:2] = rng.uniform(-3, 3, (n, 2))
    rays[: n // 4, 0] = rng.choice(zero, n // 4)
    rays[n // 8: n // 2, 1] = rng.choice(zero, n // 2 - n // 8)
    if which == 2:   # points of the tilted plane through the origin
        rays[:, 2] = -rays[:, 0] * tilt[0] / tilt[2]
        rays[::3, 0] = rng.choice(zero, len(rays[::3]))
        rays[::3, 2] = rng.choice(zero, len(rays[::3]))
    else:
        rays[:, 2] = rng.choice(zero, n)
    d = rng.normal(0, 0.15, (n, 3))
    d[:, 2] = -1.0 if which == 1 else 1.0
    d[::5, 0] = rng.choice(zero, len(d[::5]))
    d[::7, 1] = rng.choice(zero, len(d[::7]))
    d[::10, 0:2] = rng.choice(zero, (len(d[::10]), 2))
    d[::11, 2] *= -1
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 6] = rng.choice(np.array([0.0, -0.0, 1.5, -2.0]), n)
    rays[:, 7] = 0.532
    return system, rtm.Vacuum(), rtm.Vacuum(), rays


N_RANDOM_SYSTEMS = 40


CASES = {
    "plano_convex": plano_convex,
    "plano_convex_3d": lambda rt, rtm: plano_convex(rt, rtm, n_disps=31, nphis=8),
    "doublet_ebaf11": doublet_ebaf11,
    "doublet_nlak22": doublet_nlak22,
    "relay10": relay10,
    "relay10_script": relay10_script,
    "opm": opm,
    "achromat_imaging": achromat_imaging,
    "mirrors": mirrors,
    "perfect_lens_phase": perfect_lens_phase,
    "perfect_imaging": perfect_imaging,
    "cauchy_singlet": cauchy_singlet,
    "edge_mix": edge_mix,
    "reversed_doublet": reversed_doublet,
    "invalid_inputs": invalid_inputs,
    "perfect_imaging_in_focus": perfect_imaging_in_focus,
    "retro_mirror": retro_mirror,
    "on_plane_flat": lambda rt, rtm: launched_on_plane(rt, rtm, 0),
    "on_plane_flat_back": lambda rt, rtm: launched_on_plane(rt, rtm, 1),
    "on_plane_tilted": lambda rt, rtm: launched_on_plane(rt, rtm, 2),
    "on_plane_mirror": lambda rt, rtm: launched_on_plane(rt, rtm, 3),
    "on_plane_lens": lambda rt, rtm: launched_on_plane(rt, rtm, 4),
}

# cases whose refractive indices go through np.power (Ebaf11) and are therefore only bit-reproducible on a host
# whose NumPy SIMD dispatch matches the machine that wrote the golden file (tests check this and fall back to 1e-10)
POWER_DEPENDENT = {"doublet_ebaf11", "achromat_imaging"}


# ----------------------------------------------------------------------------------------------------------
# platform-independent big batches (only + - * / sqrt: no libm), for checksum pins at sizes too big to store
# ----------------------------------------------------------------------------------------------------------
def lattice_rays(n_side: int, half_width: float, z0: float, wavelength: float, tilt=(0.0, 0.0),
                 converge: float = 0.0) -> np.ndarray:
    """
    n_side**2 rays starting on a square lattice in the plane z = z0, directions
    (tx + converge*x, ty + converge*y, sqrt(1 - .^2 - .^2)); index = iy * n_side + ix.
    """
    i = np.arange(n_side, dtype=np.float64)
    step = (2.0 * half_width) / (n_side - 1)
    c = i * step - half_width
    x = np.tile(c, n_side)
    y = np.repeat(c, n_side)
    rays = np.zeros((n_side * n_side, 8))
    rays[:, 0] = x
    rays[:, 1] = y
    rays[:, 2] = z0
    dx = tilt[0] + converge * x
    dy = tilt[1] + converge * y
    rays[:, 3] = dx
    rays[:, 4] = dy
    rays[:, 5] = np.sqrt((1.0 - dx * dx) - dy * dy)
    rays[:, 7] = wavelength
    return rays
