"""
The ctypes stub of INTEGRATION.md section 2, verbatim: what a maintainer of the reference would paste in place of
System.ray_trace (raytrace.py:641-661) to call librtb.so directly.  Uses nothing from ray_trace_pb_b200 except the
built library file; surfaces / materials are read through the attributes the reference's classes have.
"""
import ctypes as C
from pathlib import Path

import numpy as np

L = C.CDLL(str(Path(__file__).resolve().parent.parent / "ray_trace_pb_b200" / "_lib" / "librtb.so"))


class RtbSurface(C.Structure):       # include/rtb.h: rtb_surface
    _fields_ = [("kind", C.c_int32), ("hints", C.c_int32), ("center", C.c_double * 3), ("normal", C.c_double * 3),
                ("input_axis", C.c_double * 3), ("radius", C.c_double), ("radius_sq", C.c_double),
                ("abs_radius", C.c_double), ("aperture_rad", C.c_double), ("focal_len", C.c_double),
                ("normal_f", C.c_double * 3), ("sin_alpha", C.c_double)]


class RtbMaterial(C.Structure):      # rtb_material
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("b", C.c_double * 3), ("c", C.c_double * 3),
                ("n_const", C.c_double)]


class RtbSystem(C.Structure):        # rtb_system
    _fields_ = [("n_surfaces", C.c_int32), ("n_wavelengths", C.c_int32), ("surfaces", C.POINTER(RtbSurface)),
                ("materials", C.POINTER(RtbMaterial)), ("wavelengths", C.POINTER(C.c_double)),
                ("n_table", C.POINTER(C.c_double))]


class RtbTraceOpts(C.Structure):     # rtb_trace_opts
    _fields_ = [("precision", C.c_int32), ("keep_mode", C.c_int32), ("n_keep", C.c_int32), ("flags", C.c_int32),
                ("keep_slabs", C.POINTER(C.c_int32)), ("reduce", C.c_void_p)]


L.rtb_trace_host.argtypes = [C.POINTER(RtbSystem), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(RtbTraceOpts), C.c_int]
L.rtb_last_error.restype = C.c_char_p

KIND = {"FlatSurface": 0, "SphericalSurface": 1, "PlaneMirror": 2, "PerfectLens": 3}


def v3(x):
    return (C.c_double * 3)(*np.asarray(x, dtype=float).reshape(3))


def pack_surface(s):
    r = RtbSurface(kind=KIND[type(s).__name__], center=v3(s.center), input_axis=v3(s.input_axis),
                   aperture_rad=s.aperture_rad)
    if r.kind == 1:
        r.normal, r.radius, r.radius_sq, r.abs_radius = v3(s.input_axis), s.radius, s.radius**2, abs(s.radius)
    else:
        r.normal = v3(s.normal)
    if r.kind == 3:
        r.focal_len, r.normal_f, r.sin_alpha = s.focal_len, v3(np.asarray(s.normal) * s.focal_len), np.sin(s.alpha)
    return r


def pack(self, rays, initial_material, final_material):
    """the host side of the call: the reference's objects -> the C structs (everything but the device call)"""
    materials = [initial_material] + self.materials + [final_material]
    if len(materials) != len(self.surfaces) + 1:
        raise ValueError("length of materials should be len(surfaces) + 1")
    rays = np.ascontiguousarray(np.atleast_2d(rays), dtype=float)                 # (N, 8)
    S = len(self.surfaces)
    wl = np.unique(rays[~np.isnan(rays[:, 7]), 7])                                # <= RTB_MAX_WAVELENGTHS (8) values
    with np.errstate(all="ignore"):                                               # (len(wl)+1, S+1): the user's own n()
        table = np.array([np.asarray(m.n(np.append(wl, np.nan)), dtype=float).reshape(-1) for m in materials]).T.copy()
    surf = (RtbSurface * S)(*map(pack_surface, self.surfaces))
    mats = (RtbMaterial * (S + 1))(*[RtbMaterial(kind=2) for _ in materials])     # 2 = "host table only": always valid
    sysd = RtbSystem(S, len(wl), surf, mats, wl.ctypes.data_as(C.POINTER(C.c_double)),
                     table.ctypes.data_as(C.POINTER(C.c_double)))
    return sysd, rays, (surf, mats, wl, table)                                    # (the tuple keeps the memory alive)


def ray_trace(self, rays, initial_material, final_material):        # drop-in for raytrace.py:641-661
    sysd, rays, _keep = pack(self, rays, initial_material, final_material)
    S, N = len(self.surfaces), rays.shape[0]
    out = np.empty((2 * S + 1, N, 8))
    opts = RtbTraceOpts(precision=0, keep_mode=0)                                 # fp64 exact, full history
    if L.rtb_trace_host(C.byref(sysd), rays.ctypes.data, N, out.ctypes.data, C.byref(opts), 0) != 0:
        raise ValueError(L.rtb_last_error().decode())
    return out
