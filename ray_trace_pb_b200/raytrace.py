"""
Drop-in host API for ``raytrace.raytrace`` (``rt``) of QI2lab/ray_trace_pb, backed by the B200 library.

A ray is the 8-vector ``(xo, yo, zo, dx, dy, dz, phase, wavelength)``: a point, a unit direction, the accumulated
phase ``2*pi/wavelength * OPL`` in radians and the wavelength (um) -- reference ``raytrace.py:1-13``.  z is the
optical axis, x is "up".  Invalid rays are marked in-band with NaN.

What runs where
---------------
* ``System.ray_trace`` / ``Surface.propagate`` -- the hot path -- pack the prescription into the POD records of
  ``include/rtb.h`` and run the fused sm_100a kernel (one thread carries one ray through every surface).  Results are
  bit-identical to the reference's NumPy path in the default fp64 mode.  There is no CPU fallback: without the
  built library or without a GPU these calls raise.
* ``intersect_rays`` and ``propagate_ray2plane`` run small device kernels of the same library.
* Everything that is O(number of surfaces) scalar algebra (ABCD matrices, cardinal points, Seidel sums, auto-focus
  bookkeeping, system editing) stays on the host, as in the reference, and keeps the reference's conventions so
  the prescriptions it builds are the ones the reference would build.

Extensions over the reference signature (all keyword-only, defaults reproduce the reference):
``System.ray_trace(..., keep="all" | "last" | [slab indices], precision="f64" | "f64_fast" | "f32", device=0)``.
"""
from __future__ import annotations

from collections.abc import Sequence
from copy import deepcopy
from typing import Optional

import numpy as np

from . import _ffi, engine
from .materials import Material, Vacuum

__all__ = ["get_free_space_abcd", "get_ray_fan", "get_collimated_rays", "intersect_rays", "propagate_ray2plane",
           "ray_angle_about_axis", "dist_pt2plane", "System", "Doublet", "Surface", "RefractingSurface",
           "ReflectingSurface", "FlatSurface", "PlaneMirror", "SphericalSurface", "PerfectLens"]


# ======================================================================================================
# small helpers
# ======================================================================================================
def get_free_space_abcd(d: float, n: float = 1.) -> np.ndarray:
    """Ray-transfer matrix of a gap of length ``d`` in index ``n``, acting on (h, n*u). Reference raytrace.py:32-41."""
    return np.array([[1, d / n], [0, 1]])


def _transverse_basis(axis: np.ndarray, fallback: bool):
    """
    Unit vectors (e1, e2) with e1 = y_hat x axis (normalised) and e2 = axis x e1 -- reference raytrace.py:79-81 and
    135-144.  ``fallback`` selects the collimated-generator behaviour (axis || y handled, e2 re-normalised).
    """
    y_hat = np.array([0, 1, 0])
    e1 = np.cross(y_hat, axis)
    if fallback and np.linalg.norm(e1) == 0:
        e1 = np.cross(axis, np.array([1, 0, 0]))
    e1 = e1 / np.linalg.norm(e1)
    e2 = np.cross(axis, e1)
    if fallback:
        e2 = e2 / np.linalg.norm(e2)
    return e1, e2


def get_ray_fan(pt, theta_max: float, n_thetas: int, wavelengths, nphis: int = 1, center_ray=(0, 0, 1)) -> np.ndarray:
    """
    Fan of ``n_thetas * nphis`` rays leaving ``pt``: polar angle theta in linspace(-theta_max, theta_max) about
    ``center_ray``, azimuth phi = 2*pi*k/nphis; row index = i_phi * n_thetas + i_theta.  Reference raytrace.py:45-96.
    (The device-side equivalent for huge fans is ``ray_trace_pb_b200.device.RaySource.fan``.)
    """
    center_ray = np.array(center_ray)
    if np.linalg.norm(center_ray) != 1:
        raise ValueError("center_ray must be a unit vector")
    theta = np.linspace(-theta_max, theta_max, n_thetas)
    phi = np.arange(nphis) * 2 * np.pi / nphis
    tt, pp = np.meshgrid(theta, phi)
    tt, pp = tt.ravel(), pp.ravel()
    e1, e2 = _transverse_basis(center_ray, fallback=False)
    pt = np.array(pt).squeeze()

    rays = np.zeros((n_thetas * nphis, 8))
    rays[:, 0], rays[:, 1], rays[:, 2] = pt[0], pt[1], pt[2]
    cos_t, sin_t, cos_p, sin_p = np.cos(tt), np.sin(tt), np.cos(pp), np.sin(pp)
    for k in range(3):
        rays[:, 3 + k] = center_ray[k] * cos_t + e1[k] * cos_p * sin_t + e2[k] * sin_p * sin_t
    rays[:, 6] = 0
    rays[:, 7] = wavelengths
    return rays


def get_collimated_rays(pt, displacement_max, n_disps: int, wavelengths, nphis: int = 1, phi_start: float = 0.,
                        normal=(0, 0, 1)) -> np.ndarray:
    """
    Bundle of parallel rays along ``normal`` starting in the plane through ``pt`` orthogonal to it: radial offsets
    linspace(-displacement_max, displacement_max, n_disps), azimuths 2*pi*k/nphis + phi_start; row index =
    i_disp * nphis + i_phi.  Reference raytrace.py:99-161.
    """
    if np.abs(np.linalg.norm(normal) - 1) > 1e-12:
        raise ValueError("normal must be a normalized vector")
    phi = np.arange(nphis) * 2 * np.pi / nphis + phi_start
    off = np.linspace(-displacement_max, displacement_max, n_disps)
    pp, oo = np.meshgrid(phi, off)
    pp, oo = pp.ravel(), oo.ravel()
    pt = np.array(pt).squeeze()
    normal = np.array(normal).squeeze()
    e1, e2 = _transverse_basis(normal, fallback=True)

    rays = np.zeros((n_disps * nphis, 8))
    rays[:, 0:3] = pt[None, :] + e1[None, :] * (oo * np.cos(pp))[:, None] + e2[None, :] * (oo * np.sin(pp))[:, None]
    rays[:, 3], rays[:, 4], rays[:, 5] = normal[0], normal[1], normal[2]
    rays[:, 6] = 0
    rays[:, 7] = wavelengths
    return rays


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def intersect_rays(ray1, ray2):
    """
    Point where pairs of rays meet in free space (NaN where they do not meet to within 1e-12, or are parallel).
    One of the inputs may be a single ray.  Runs on the device (kernel ``intersect_kernel``); NumPy in -> NumPy out,
    CUDA tensors in -> CUDA tensor out.  Reference raytrace.py:164-238.
    """
    from . import device as dev
    return dev.intersect_rays(ray1, ray2)


def propagate_ray2plane(rays, normal, center, material: Material, exclude_backward_propagation: bool = False):
    """
    Intersect rays with the plane(s) ``(p - center) . normal = 0``; returns ``(rays_out, ts)``.  ``normal`` and
    ``center`` broadcast to (N, 3).  Implemented as a one-surface device trace per distinct plane is not possible for
    per-ray planes, so this helper runs the dedicated ``ray2plane`` device kernel.  Reference raytrace.py:241-306.
    """
    from . import device as dev
    return dev.propagate_ray2plane(rays, normal, center, material, exclude_backward_propagation)


def ray_angle_about_axis(rays, reference_axis):
    """Angle of each ray to ``reference_axis`` and the unit vector of its transverse part. Reference raytrace.py:309-328."""
    rays = np.atleast_2d(rays)
    axis = np.asarray(reference_axis)
    d = rays[:, 3:6]
    cosines = np.sum(d * axis[None, :], axis=1)
    transverse = d - cosines[:, None] * axis[None, :]
    transverse = transverse / np.linalg.norm(transverse, axis=1)[:, None]
    return np.arccos(cosines), transverse


def dist_pt2plane(pts, normal, center):
    """Distance from points to a plane and the nearest points on it. Reference raytrace.py:331-353."""
    pts = np.atleast_2d(pts)
    npts = pts.shape[0]
    probe = np.concatenate((pts, np.tile(normal, (npts, 1)), np.zeros((npts, 2))), axis=1)
    with np.errstate(all="ignore"):
        hit, _ = propagate_ray2plane(probe, normal, center, Vacuum())
    nearest = hit[:, :3]
    return np.linalg.norm(nearest - pts, axis=1), nearest


# ======================================================================================================
# System
# ======================================================================================================
class System:
    """
    An ordered list of surfaces with the media between them (``len(materials) == len(surfaces) - 1``); the media
    before the first and after the last surface are given per call.  Reference raytrace.py:359-933.
    """

    def __init__(self, surfaces: list, materials: list, names=None, surfaces_by_name=None,
                 aperture_stop: Optional[int] = None):
        if len(materials) > 1 and len(materials) != len(surfaces) - 1:
            raise ValueError(f"len(materials) = {len(materials):d} != len(surfaces) - 1 = {len(surfaces) - 1:d}")
        self.surfaces = surfaces
        self.materials = materials
        self.aperture_stop = aperture_stop
        if names is None:
            names = [""]
        elif not isinstance(names, list):
            names = [names]
        self.names = names
        if surfaces_by_name is None:
            self.surfaces_by_name = np.zeros(len(surfaces), dtype=int)
        else:
            if len(surfaces_by_name) != len(surfaces):
                raise ValueError("len(surfaces_by_name) must equal len(surfaces)")
            self.surfaces_by_name = np.array(surfaces_by_name).astype(int)

    # ------------------------------------------------------------------ the hot path
    def ray_trace(self, rays, initial_material: Material, final_material: Material, *, keep="all",
                  precision: str = "f64", device: int = 0):
        """
        Trace ``rays`` through every surface on the GPU.

        rays: ``(8,)``, ``(N, 8)`` or an existing history ``(K, N, 8)`` (its last slab is traced and the result is
        appended).  Returns ``(K + 2*S, N, 8)``: slab 0 the launch rays, slab 2k+1 the rays at surface k, slab 2k+2
        just after it -- exactly the array the reference returns (raytrace.py:641-661, 1229-1232).

        ``keep="last"`` or a list of slab indices returns only those slabs of the new trace (shape
        ``(n_kept, N, 8)``), which is what makes 1e8-ray batches fit.  A CUDA ``torch.Tensor`` input stays on the
        device and a tensor is returned.
        """
        materials = [initial_material] + list(self.materials) + [final_material]
        if len(materials) != len(self.surfaces) + 1:
            raise ValueError("length of materials should be len(surfaces) + 1")
        if _is_torch_cuda(rays):
            from . import device as dev
            return dev.trace_tensor(self.surfaces, materials, rays, keep=keep, precision=precision)

        rays = np.asarray(rays, dtype=np.float64)
        if rays.ndim == 1:
            rays = rays[None, None, :]
        elif rays.ndim == 2:
            rays = rays[None, :, :]
        if rays.ndim != 3 or rays.shape[-1] != 8:
            raise ValueError(f"rays must have shape (8,), (N, 8) or (K, N, 8), got {rays.shape}")
        prior, launch = rays[:-1], rays[-1]
        if isinstance(keep, str) and keep == "all":
            n_new = 2 * len(self.surfaces) + 1
            out = _ffi.result_array((prior.shape[0] + n_new, launch.shape[0], 8))
            out[:prior.shape[0]] = prior
            engine.trace_host(self.surfaces, materials, launch, keep="all", precision=precision, device=device,
                              out=out[prior.shape[0]:])
            return out
        return engine.trace_host(self.surfaces, materials, launch, keep=keep, precision=precision, device=device)

    # ------------------------------------------------------------------ editing
    def reverse(self):
        """
        The same optic seen from the other side: surfaces and media in reverse order, every surface's input and
        output axis negated (geometry untouched).  Names are dropped.  Reference raytrace.py:402-415.
        """
        flipped = []
        for s in reversed(self.surfaces):
            s = deepcopy(s)
            s.input_axis *= -1
            s.output_axis *= -1
            flipped.append(s)
        return System(flipped, list(reversed(self.materials)))

    def concatenate(self, other, material: Material, distance: Optional[float] = None,
                    axis: Sequence[float] = (0., 0., 1.)):
        """
        New system = this one, then ``material``, then ``other`` (a System or a single Surface; deep-copied).
        With ``distance`` the appended part is translated so that its first paraxial centre sits ``distance`` along
        ``axis`` after this system's last paraxial centre, keeping its internal spacing.  Reference raytrace.py:417-478.
        """
        if isinstance(other, System):
            added = [deepcopy(s) for s in other.surfaces]
            added_materials = other.materials
            added_stop = other.aperture_stop
            added_by_name = other.surfaces_by_name
            added_names = other.names
            originals = other.surfaces
        elif isinstance(other, Surface):
            added = [deepcopy(other)]
            added_materials = []
            added_stop = None
            added_by_name = np.array([0])
            added_names = [""]
            originals = [other]
        else:
            raise TypeError(f"other should be of type System or Surface, but was {type(other)}")

        if distance is not None:
            for k, s in enumerate(added):
                if k == 0:
                    shift = self.surfaces[-1].paraxial_center + distance * np.array(axis) - s.paraxial_center
                else:
                    # keep the spacing to the previous surface: new_k = new_{k-1} + (old_k - old_{k-1})
                    shift = added[k - 1].paraxial_center - originals[k - 1].paraxial_center
                s.center += shift
                s.paraxial_center += shift

        by_name = np.concatenate((self.surfaces_by_name, added_by_name + np.max(self.surfaces_by_name) + 1))
        stop = self.aperture_stop
        if stop is None and added_stop is not None:
            stop = added_stop + len(self.surfaces)
        return System(self.surfaces + added, self.materials + [material] + added_materials,
                      names=self.names + added_names, surfaces_by_name=by_name, aperture_stop=stop)

    def set_aperture_stop(self, surface_index: int):
        self.aperture_stop = surface_index

    # ------------------------------------------------------------------ paraxial analysis (host, O(S) scalars)
    def _indices(self, wavelength, initial_material, final_material):
        media = [initial_material] + self.materials + [final_material]
        return np.array([m.n(wavelength) for m in media])

    def get_ray_transfer_matrix(self, wavelength: float, initial_material: Material, final_material: Material,
                                axis=None) -> np.ndarray:
        """
        Cumulative ABCD matrices, shape (S+1, 2, 2): entry i maps a launch ray (h, n*u) at the first surface to just
        before surface i; the last entry to just after the last surface.  Gaps are measured between paraxial
        centres.  Reference raytrace.py:719-752.
        """
        ns = self._indices(wavelength, initial_material, final_material)
        S = len(self.surfaces)
        mats = np.zeros((S + 1, 2, 2))
        mats[0] = get_free_space_abcd(0, ns[0])
        for i in range(1, S + 1):
            refract = self.surfaces[i - 1].get_ray_transfer_matrix(ns[i - 1], ns[i])
            if i < S:
                gap = np.linalg.norm(self.surfaces[i].paraxial_center - self.surfaces[i - 1].paraxial_center)
                step = get_free_space_abcd(gap, ns[i]).dot(refract)
            else:
                step = refract
            mats[i] = step.dot(mats[i - 1])
        return mats

    def get_cardinal_points(self, wavelength: float, initial_material: Material, final_material: Material, axis=None):
        """
        ``fp1, fp2, pp1, pp2, np1, np2, efl1, efl2``: focal, principal and nodal points (3-vectors on the axis of the
        first / last surface) and the two effective focal lengths.  Reference raytrace.py:754-813.
        """
        fwd = self.get_ray_transfer_matrix(wavelength, initial_material, final_material)[-1]
        bwd = self.reverse().get_ray_transfer_matrix(wavelength, final_material, initial_material)[-1]
        n_obj = initial_material.n(wavelength)
        n_img = final_material.n(wavelength)
        first, last = self.surfaces[0], self.surfaces[-1]

        # image side: a gap d2 after the last surface zeroes the A element -> back focal point
        d2 = -fwd[0, 0] / fwd[1, 0] * n_img
        efl2 = -n_img / fwd[1, 0]
        fp2 = last.paraxial_center + d2 * last.output_axis
        pp2 = fp2 - efl2 * last.output_axis
        d2_nodal = (n_img - n_obj * bwd[1, 1]) / bwd[1, 0]
        np2 = last.paraxial_center + d2_nodal * last.output_axis

        # object side: same construction on the reversed system
        d1 = -bwd[0, 0] / bwd[1, 0] * n_obj
        efl1 = -n_obj / bwd[1, 0]
        fp1 = first.paraxial_center - d1 * first.input_axis
        pp1 = fp1 + efl1 * first.input_axis
        d1_nodal = (n_obj - n_img * fwd[1, 1]) / fwd[1, 0]
        np1 = first.paraxial_center - d1_nodal * first.output_axis
        return fp1, fp2, pp1, pp2, np1, np2, efl1, efl2

    def find_paraxial_collimated_distance(self, other, wavelength: float, initial_material: Material,
                                          intermediate_material: Material, final_material: Material, axis=None):
        """Gap between this system and ``other`` that makes the pair afocal. Reference raytrace.py:615-639."""
        m1 = self.get_ray_transfer_matrix(wavelength, initial_material, intermediate_material)[-1]
        m2 = other.get_ray_transfer_matrix(wavelength, intermediate_material, final_material)[-1]
        return -(m1[0, 0] / m1[1, 0] + m2[1, 1] / m2[1, 0]) * intermediate_material.n(wavelength)

    def gaussian_paraxial(self, q_in: complex, wavelength: float, initial_material: Material,
                          final_material: Material, print_results: bool = False):
        """Complex beam parameter q at every surface (q' = (A q + B) / (C q + D)). Reference raytrace.py:663-717."""
        S = len(self.surfaces)
        ns = np.zeros(S + 1)
        qs = np.zeros(S + 1, dtype=complex)
        qs[0] = q_in
        for i, s in enumerate(self.surfaces):
            n1 = initial_material.n(wavelength) if i == 0 else self.materials[i - 1].n(wavelength)
            if i < S - 1:
                n2 = self.materials[i].n(wavelength)
                gap = np.linalg.norm(self.surfaces[i + 1].paraxial_center - s.paraxial_center)
            else:
                n2 = final_material.n(wavelength)
                gap = 0.
            m = get_free_space_abcd(gap, n2).dot(s.get_ray_transfer_matrix(n1, n2))
            qs[i + 1] = (qs[i] * m[0, 0] + m[0, 1]) / (qs[i] * m[1, 0] + m[1, 1])
            ns[i], ns[i + 1] = n1, n2
        if print_results:
            # beam radius / waist from q: 1/q = 1/R - i*lambda/(pi n w^2)
            print("surface          R           w          wo           z          zr")
            for i in range(S + 1):
                inv = 1 / qs[i]
                lam = wavelength / ns[i]
                with np.errstate(divide="ignore"):
                    radius = np.inf if inv.real == 0 else 1 / inv.real
                    w = np.sqrt(-lam / (np.pi * inv.imag))
                z, zr = qs[i].real, qs[i].imag
                wo = np.sqrt(lam * zr / np.pi)
                print(f"{i:02d}: {radius:10.6g}, {w:10.6g}, {wo:10.6g}, {z:10.6g}, {zr:10.6g}")
        return qs

    def seidel_third_order(self, wavelength: float, initial_material: Material, final_material: Material,
                           print_results: bool = False, object_distance: float = 0., object_height: float = 0.,
                           object_angle: float = 0.):
        """
        Third-order Seidel sums per surface, columns [spherical, coma, astigmatism, field curvature, distortion]
        (Kidger, "Fundamental Optical Design", eqs. 6.27-6.30, 6.37).  The object sits ``object_distance`` before the
        first surface (``np.inf`` -> collimated input with field angle ``object_angle``).  Needs an aperture stop.
        Reference raytrace.py:484-613.
        """
        if self.aperture_stop is None:
            raise ValueError("aperture_stop was None, but aperture_stop must be provided to "
                             "compute Seidel aberrations")
        ns = self._indices(wavelength, initial_material, final_material)
        mats = self.get_ray_transfer_matrix(wavelength, initial_material, final_material)
        to_stop = mats[self.aperture_stop]
        stop_radius = self.surfaces[self.aperture_stop].aperture_rad

        # marginal ray (fills the stop) and chief ray (through the stop centre) at the first surface, as (h, u)
        if np.isinf(object_distance):
            h_m, u_m = stop_radius / to_stop[0, 0], 0.
            h_c, u_c = 0., object_angle
        else:
            obj2stop = to_stop.dot(get_free_space_abcd(object_distance, ns[0]))
            a, b, c, d = obj2stop[0, 0], obj2stop[0, 1], obj2stop[1, 0], obj2stop[1, 1]
            u0_m = stop_radius / b / ns[0]
            h_m = a * 0. + b * ns[0] * u0_m
            u_m = c * 0. + d * ns[0] * u0_m
            u0_c = -a / b / ns[0] * object_height
            h_c = a * object_height + b * ns[0] * u0_c
            u_c = c * object_height + d * ns[0] * u0_c

        launch = np.array([[h_m, h_c], [ns[0] * u_m, ns[0] * u_c]])
        traced = mats.dot(launch)                      # (S+1, 2, 2): [:, 0, :] heights, [:, 1, :] n*u; [..., 0] marginal
        h, hbar = traced[:-1, 0, 0], traced[:-1, 0, 1]
        nu, nubar = traced[:-1, 1, 0], traced[:-1, 1, 1]
        n_before, n_after = ns[:-1], ns[1:]

        curv = np.array([1 / s.radius if isinstance(s, SphericalSurface) else 0 for s in self.surfaces])
        A = n_before * h * curv + nu                   # refraction invariant, marginal
        Abar = n_before * hbar * curv + nubar          # refraction invariant, chief
        delta_un = traced[1:, 1, 0] / n_after / n_after - nu / n_before / n_before
        lagrange = n_before * (hbar * nu / n_before - h * nubar / n_before)

        ab = np.zeros((len(self.surfaces), 5)) * np.nan
        ab[:, 0] = -A**2 * h * delta_un
        ab[:, 1] = -A * Abar * h * delta_un
        ab[:, 2] = -Abar ** 2 * h * delta_un
        ab[:, 3] = -lagrange ** 2 * curv * (1 / n_after - 1 / n_before)
        ab[:, 4] = (-Abar ** 3 * h * (1 / n_after**2 - 1 / n_before**2) +
                    hbar * Abar * curv * (2 * h * Abar - hbar * A) * (1 / n_after - 1 / n_before))

        if print_results:
            print("surface,          h,          u,       hbar,       ubar,   delta(u/n)          A,       Abar,   Lag. inv.")
            for i in range(len(self.surfaces)):
                print(f"{i:02d}:      {h[i]:10.6g}, {nu[i] / ns[i]:10.6g}, {hbar[i]:10.6g}, {nubar[i] / ns[i]:10.6g}, "
                      f"{delta_un[i]:10.6g}, {A[i]:10.6g}, {Abar[i]:10.6g}, {lagrange[i]:10.6g}")
            print("surfaces, spherical,       coma,     astig.,   field curv.,   distortion")
            for i in range(len(self.surfaces)):
                print(f"{i:02d}:      " + ", ".join(f"{v:10.6g}" for v in ab[i]))
            print("sum:     " + ", ".join(f"{v:10.6g}" for v in np.sum(ab, axis=0)))
        return ab

    def auto_focus(self, wavelength: float, initial_material: Material, final_material: Material,
                   mode: str = "ray-fan"):
        """
        Focus after the last surface: "ray-fan" / "collimated" trace three nearly-paraxial rays (on the GPU) and
        intersect the outer two; "paraxial-focused" / "paraxial-collimated" use the ABCD matrix.
        Reference raytrace.py:815-855.
        """
        if mode in ("ray-fan", "collimated"):
            if mode == "ray-fan":
                probe = get_ray_fan([0, 0, 0], 1e-9, 3, wavelength)
            else:
                probe = get_collimated_rays([0, 0, 0], 1e-9, 3, wavelength)
            traced = self.ray_trace(probe, initial_material, final_material)
            return intersect_rays(traced[-1, 1], traced[-1, 2])[0]
        if mode == "paraxial-focused":
            return self.get_cardinal_points(wavelength, initial_material, final_material)[1]
        if mode == "paraxial-collimated":
            m = self.get_ray_transfer_matrix(wavelength, initial_material, final_material)[-1]
            dx = -m[0, 0] / m[1, 0] * self.materials[-1].n(wavelength)
            last = self.surfaces[-1]
            return last.paraxial_center[2] + dx * np.sign(last.input_axis[2])
        raise ValueError(f"mode must be 'ray-fan', or 'collimated' 'paraxial-focused',"
                         f" or paraxial-collimated' but was '{mode:s}'")

    # ------------------------------------------------------------------ drawing (matplotlib imported lazily)
    def plot(self, ray_array=None, phi: float = 0, colors=None, label: str = None, ax=None,
             show_names: bool = True, fontsize: float = 16, **kwargs):
        """Side view (z horizontal, height in the azimuthal plane ``phi``) of rays and surfaces. Reference raytrace.py:857-932."""
        import matplotlib.pyplot as plt
        if ax is None:
            fig = plt.figure(**kwargs)
            ax = plt.subplot(1, 1, 1)
        else:
            fig = ax.get_figure()
        if ray_array is not None:
            ray_array = np.asarray(ray_array)
            height = ray_array[:, :, 0] * np.cos(phi) + ray_array[:, :, 1] * np.sin(phi)
            label = "" if label is None else label
            if colors is None:
                ax.plot(ray_array[:, :, 2], height, label=label)
            else:
                if len(colors) == 1 and not isinstance(colors, list):
                    colors = [colors] * ray_array.shape[1]
                if len(colors) != ray_array.shape[1]:
                    raise ValueError("len(colors) must equal ray_array.shape[1]")
                for k in range(ray_array.shape[1]):
                    ax.plot(ray_array[:, k, 2], height[:, k], color=colors[k], **({"label": label} if k == 0 else {}))
            ax.set_xlabel("z-position (mm)", fontsize=fontsize)
            ax.set_ylabel("height (mm)", fontsize=fontsize)
        ax.tick_params(axis="x", labelsize=fontsize)
        ax.tick_params(axis="y", labelsize=fontsize)
        for k, s in enumerate(self.surfaces or []):
            s.draw(ax)
            if show_names and (k == 0 or self.surfaces_by_name[k] != self.surfaces_by_name[k - 1]):
                ax.text(s.paraxial_center[2], s.paraxial_center[0] + 1.1 * s.aperture_rad,
                        self.names[self.surfaces_by_name[k]], horizontalalignment="center", fontsize=fontsize)
        return fig, ax


class Doublet(System):
    """
    Cemented doublet from catalogue data.  Radii are quoted crown-side-left (positive = convex towards -z);
    ``input_collimated=True`` puts the crown first, ``False`` mounts the lens flipped (flint first, radii negated).
    An infinite radius gives a flat face.  Reference raytrace.py:935-1025.
    """

    def __init__(self, material_crown=None, material_flint=None, radius_crown=None, radius_flint=None,
                 radius_interface=None, thickness_crown=None, thickness_flint=None, aperture_radius: float = 25.4,
                 input_collimated: bool = True, names: str = ""):
        if input_collimated:
            media = [material_crown, material_flint]
            radii = [radius_crown, radius_interface, radius_flint]
            zs = [0, thickness_crown, thickness_crown + thickness_flint]
        else:
            media = [material_flint, material_crown]
            radii = [-radius_flint, -radius_interface, -radius_crown]
            zs = [0, thickness_flint, thickness_flint + thickness_crown]
        faces = []
        for r, z in zip(radii, zs):
            if np.isinf(r):
                faces.append(FlatSurface([0, 0, z], [0, 0, 1], aperture_rad=aperture_radius))
            else:
                faces.append(SphericalSurface.get_on_axis(r, z, aperture_radius))
        self.radius_crown = float(radius_crown)
        self.radius_flint = float(radius_flint)
        self.radius_interface = float(radius_interface)
        self.thickness_crown = float(thickness_crown)
        self.thickness_flint = float(thickness_flint)
        super().__init__(faces, media, names=names, surfaces_by_name=None)


# ======================================================================================================
# Surfaces
# ======================================================================================================
class Surface:
    """
    Base class.  Geometry is held as ``center`` (a point defining the surface), ``paraxial_center`` (where the
    optical axis pierces it), ``input_axis`` / ``output_axis`` (direction of travel before / after) and
    ``aperture_rad``.  Reference raytrace.py:1031-1156.

    Subclasses that can be traced provide ``device_record()`` (the numbers the kernel needs).
    """

    def __init__(self, input_axis, output_axis, center, paraxial_center, aperture_rad: float):
        self.input_axis = np.array(input_axis).squeeze().astype(float)
        self.output_axis = np.array(output_axis).squeeze().astype(float)
        self.center = np.array(center).squeeze().astype(float)
        self.paraxial_center = np.array(paraxial_center).squeeze().astype(float)
        self.aperture_rad = aperture_rad

    # -- the per-surface operator: a one-surface trace on the device ------------------------------------
    def propagate(self, ray_array, material1: Material, material2: Material = None):
        """
        Append the two slabs "at this surface" and "just after it" to ``ray_array`` ((8,), (N, 8) or (K, N, 8)).
        Reference raytrace.py:1160-1234 / 1238-1303 / 1601-1801.
        """
        if material2 is None:
            material2 = material1
        rays = np.asarray(ray_array, dtype=np.float64)
        if rays.ndim == 1:
            rays = rays[None, None, :]
        elif rays.ndim == 2:
            rays = rays[None, :, :]
        both = engine.trace_host([self], [material1, material2], rays[-1], keep=[1, 2])
        return np.concatenate((rays, both), axis=0)

    def get_intersect(self, rays, material: Material):
        """Rays advanced to this surface (phase included), before any cull of ``propagate``. Reference raytrace.py:1081-1090."""
        from . import device as dev
        return dev.surface_intersect(self, rays, material)

    def get_normal(self, pts):
        raise NotImplementedError

    def is_pt_on_surface(self, pts):
        raise NotImplementedError

    def get_ray_transfer_matrix(self, n1: float, n2: float):
        raise NotImplementedError

    def solve_img_eqn(self, s, n1: float, n2: float):
        """
        Image distance for object distance ``s`` (both negative to the left of the surface), from the B = 0 condition
        of gap * surface * gap.  Reference raytrace.py:1115-1138.
        """
        m = self.get_ray_transfer_matrix(n1, n2)
        with np.errstate(divide="ignore"):
            if np.abs(s) > 1e12:
                return np.atleast_1d(-n2 * m[0, 0] / m[1, 0])
            return np.atleast_1d(-n2 * (-m[0, 0] * s / n1 + m[0, 1]) / np.array(-m[1, 0] * s / n1 + m[1, 1]))

    def draw(self, ax):
        raise NotImplementedError

    # -- shared by the planar surfaces ---------------------------------------------------------------------
    def _draw_plane_section(self, ax, infinite_ok: bool):
        y_hat = np.array([0, 1, 0])
        n_xz = self.normal - self.normal.dot(y_hat) * y_hat
        n_xz = n_xz / np.linalg.norm(n_xz)
        along = np.cross(n_xz, y_hat)
        if infinite_ok and np.isinf(self.aperture_rad):
            p0, p1 = self.center, self.center + along
            ax.axline(p0[[2, 0]], xy2=p1[[2, 0]], color="k")
            return
        ts = np.linspace(-self.aperture_rad, self.aperture_rad, 101)
        pts = self.center[None, :] + ts[:, None] * along[None, :]
        ax.plot(pts[:, 2], pts[:, 0], "k")


class RefractingSurface(Surface):
    """Marker base of the surfaces that apply Snell's law (reference raytrace.py:1159-1234)."""


class ReflectingSurface(Surface):
    """Marker base of the surfaces that apply the law of reflection (reference raytrace.py:1237-1303)."""


class _PlaneGeometry:
    """get_normal / is_pt_on_surface shared by FlatSurface, PlaneMirror (host-side convenience, O(N) NumPy)."""

    def get_normal(self, pts):
        pts = np.atleast_2d(pts)
        return np.tile(np.atleast_2d(self.normal), (pts.shape[0], 1))

    def is_pt_on_surface(self, pts):
        pts = np.atleast_2d(pts)
        rel = pts[..., 0:3] - self.center
        with np.errstate(invalid="ignore"):
            on_plane = np.abs(np.sum(rel * self.normal, axis=-1)) < 1e-12
            inside = np.linalg.norm(rel, axis=-1) <= self.aperture_rad
        return np.logical_and(on_plane, inside)


class FlatSurface(_PlaneGeometry, RefractingSurface):
    """Plane ``(p - center) . normal = 0`` with a circular aperture; ``normal`` points along the direction of travel.
    Reference raytrace.py:1306-1374."""

    def __init__(self, center, normal, aperture_rad: float):
        self.normal = np.array(normal).squeeze()
        super().__init__(normal, normal, center, center, aperture_rad)

    def get_ray_transfer_matrix(self, n1=None, n2=None):
        return np.array([[1, 0], [0, 1]])

    def device_record(self):
        return {"kind": _ffi.SURF_FLAT, "center": self.center, "normal": self.normal,
                "input_axis": self.input_axis, "aperture_rad": float(self.aperture_rad)}

    def draw(self, ax):
        self._draw_plane_section(ax, infinite_ok=True)


class PlaneMirror(_PlaneGeometry, ReflectingSurface):
    """Plane mirror with a circular aperture. Reference raytrace.py:1377-1432."""

    def __init__(self, center, normal, aperture_rad):
        self.normal = np.array(normal).squeeze()
        super().__init__(normal, normal, center, center, aperture_rad)

    def get_ray_transfer_matrix(self, n1: float = None, n2: float = None):
        return np.array([[1, 0], [0, -1]])

    def device_record(self):
        return {"kind": _ffi.SURF_MIRROR, "center": self.center, "normal": self.normal,
                "input_axis": self.input_axis, "aperture_rad": float(self.aperture_rad)}

    def draw(self, ax):
        self._draw_plane_section(ax, infinite_ok=False)


class SphericalSurface(RefractingSurface):
    """
    Sphere of signed ``radius`` about ``center``; the vertex (paraxial centre) is ``center - radius * input_axis``,
    so a positive radius is convex towards the incoming light.  Reference raytrace.py:1435-1555.
    """

    def __init__(self, radius, center, aperture_rad, input_axis=(0, 0, 1)):
        self.radius = radius
        vertex = np.array(center).squeeze() - self.radius * np.array(input_axis).squeeze()
        super().__init__(input_axis, input_axis, center, vertex, aperture_rad)

    @classmethod
    def get_on_axis(cls, radius: float, surface_z_position: float, aperture_rad: float):
        """Sphere whose vertex sits at z = ``surface_z_position`` on the z axis."""
        return cls(radius, [0, 0, surface_z_position + radius], aperture_rad, (0, 0, 1))

    def get_normal(self, pts):
        pts = np.atleast_2d(pts)[:, :3]
        return (pts - np.asarray(self.center)[None, :]) / self.radius

    def is_pt_on_surface(self, pts):
        pts = np.atleast_2d(pts)
        p = pts[..., 0:3]
        with np.errstate(invalid="ignore"):
            on_sphere = np.abs(np.linalg.norm(p - self.center, axis=-1) - abs(self.radius)) < 1e-12
            off_axis = p - np.sum(p * self.input_axis, axis=-1)[..., None] * self.input_axis
            inside = np.linalg.norm(off_axis, axis=-1) <= self.aperture_rad
        return np.logical_and(on_sphere, inside)

    def get_ray_transfer_matrix(self, n1: float, n2: float) -> np.ndarray:
        # sign: +1 when the centre of curvature lies downstream of the vertex
        sgn = np.sign(np.dot(self.center - self.paraxial_center, self.input_axis))
        with np.errstate(divide="ignore"):
            f = sgn * np.abs(self.radius) / np.array(n2 - n1)
        return np.array([[1, 0], [-1 / f, 1]])

    def device_record(self):
        return {"kind": _ffi.SURF_SPHERE, "center": self.center, "normal": self.input_axis,
                "input_axis": self.input_axis, "radius": float(self.radius), "radius_sq": float(self.radius ** 2),
                "abs_radius": float(abs(self.radius)), "aperture_rad": float(self.aperture_rad)}

    def draw(self, ax):
        half_angle = np.arcsin(self.aperture_rad / np.abs(self.radius))
        th = np.linspace(-half_angle, half_angle, 101)
        ax.plot(self.center[2] - self.radius * np.cos(th), self.center[0] - self.radius * np.sin(th), "k")


class PerfectLens(_PlaneGeometry, RefractingSurface):
    """
    Ideal (Abbe-sine) lens of zero thickness at ``center``: maps (height, sin(theta)) in its front focal plane to
    (n1 f sin(theta), -height / (f n2)) in its back focal plane; the focal planes sit n1*f before and n2*f after the
    lens; rays steeper than ``alpha`` on either side are dropped.  Reference raytrace.py:1558-1821.
    """

    def __init__(self, focal_len: float, center, normal, alpha: float):
        self.focal_len = focal_len
        self.alpha = alpha
        self.normal = np.array(normal).squeeze()
        super().__init__(normal, normal, center, center, focal_len * np.sin(self.alpha))

    def is_pt_on_surface(self, pts):
        pts = np.atleast_2d(pts)
        rel = pts[:, 0:3] - self.center
        with np.errstate(invalid="ignore"):
            return np.abs(rel[:, 0] * self.normal[0] + rel[:, 1] * self.normal[1] + rel[:, 2] * self.normal[2]) < 1e-12

    def get_ray_transfer_matrix(self, n1: float = None, n2: float = None) -> np.ndarray:
        return np.array([[1, 0], [-1 / self.focal_len, 1]])

    def device_record(self):
        return {"kind": _ffi.SURF_PERFECT_LENS, "center": self.center, "normal": self.normal,
                "input_axis": self.input_axis, "aperture_rad": float(self.aperture_rad),
                "focal_len": float(self.focal_len), "normal_f": np.asarray(self.normal) * self.focal_len,
                "sin_alpha": float(np.sin(self.alpha))}

    def draw(self, ax):
        self._draw_plane_section(ax, infinite_ok=False)
