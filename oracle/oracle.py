"""
oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle: a ctypes wrapper around ``rt_oracle.c`` (the scalar restatement of the reference's
``System.ray_trace``) plus NumPy restatements of the small pieces that are NumPy/libm-defined in the reference
(ray generators) and the plain definitions of this project's own reductions (spot statistics, pupil grid).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module.  It never imports anything from ``ray_trace_pb_b200``: surfaces and materials are read by duck typing
(attribute names of the reference API: ``center``, ``normal``, ``input_axis``, ``radius``, ``aperture_rad``,
``focal_len``, ``alpha``; ``material.n(wavelengths)``), so the same code accepts the reference's objects (when
generating golden vectors in the build container) and the product's host objects (in the parity tests).

Pinning status: pinned against live outputs of the reference itself, ``tests/golden/*.npz`` written by
``tests/golden/make_golden.py`` in the build container, checked bit for bit by ``tests/test_oracle_golden.py``.
(The reference's own tests do not cover this path.)
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_LIB_PATH = _BUILD / "librt_oracle.so"
_SRC = _HERE / "rt_oracle.c"

SURF_STRIDE = 20
KIND_BY_CLASS = {"FlatSurface": 0, "SphericalSurface": 1, "PlaneMirror": 2, "PerfectLens": 3}

_lib = None


def build(force: bool = False) -> Path:
    """Compile rt_oracle.c (gcc, no FMA contraction, OpenMP for the baseline leg)."""
    if (not force) and _LIB_PATH.exists() and _LIB_PATH.stat().st_mtime >= _SRC.stat().st_mtime:
        return _LIB_PATH
    _BUILD.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
           "-fexcess-precision=standard", "-fopenmp", str(_SRC), "-o", str(_LIB_PATH), "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_trace.argtypes = [dp, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int32), dp, ctypes.c_int64, dp,
                                   ctypes.c_int, ctypes.c_int]
        L.oracle_trace.restype = ctypes.c_int
        L.oracle_sellmeier.argtypes = [dp, dp, dp, ctypes.c_int64, dp]
        L.oracle_sellmeier.restype = None
        L.oracle_intersect_rays.argtypes = [dp, dp, ctypes.c_int64, dp]
        L.oracle_intersect_rays.restype = None
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------------------------------------
# packing (duck typed)
# ----------------------------------------------------------------------------------------------------------
def pack_surfaces(surfaces) -> np.ndarray:
    """Rows of SURF_STRIDE doubles as documented at the top of rt_oracle.c."""
    out = np.zeros((len(surfaces), SURF_STRIDE))
    for i, s in enumerate(surfaces):
        name = None
        for klass in type(s).__mro__:
            if klass.__name__ in KIND_BY_CLASS:
                name = klass.__name__
                break
        if name is None:
            raise NotImplementedError(f"oracle has no restatement for surface type {type(s).__name__}")
        kind = KIND_BY_CLASS[name]
        row = out[i]
        row[0] = kind
        row[1:4] = np.asarray(s.center, dtype=float)
        row[7:10] = np.asarray(s.input_axis, dtype=float)
        row[13] = s.aperture_rad
        if kind == 1:
            row[4:7] = np.asarray(s.input_axis, dtype=float)
            row[10] = s.radius
            row[11] = s.radius ** 2      # raytrace.py:1499
            row[12] = abs(s.radius)      # raytrace.py:1528
        else:
            row[4:7] = np.asarray(s.normal, dtype=float)
        if kind == 3:
            row[14] = s.focal_len
            row[15:18] = np.asarray(s.normal) * s.focal_len   # raytrace.py:1683
            row[18] = np.sin(s.alpha)                         # raytrace.py:1758
    return out


def index_table(materials, wavelengths: np.ndarray):
    """
    One table row per distinct wavelength of the batch: ``ntab[row, j] = materials[j].n(unique_wavelengths)[row]``,
    evaluated by the materials' own (reference or product) Python ``n``.  NaN wavelengths share one row.
    """
    uniq, inv = np.unique(wavelengths, return_inverse=True)
    ntab = np.empty((len(uniq), len(materials)))
    with np.errstate(all="ignore"):
        for j, m in enumerate(materials):
            ntab[:, j] = np.asarray(m.n(uniq), dtype=float).reshape(-1)
    return np.ascontiguousarray(ntab), np.ascontiguousarray(inv.reshape(-1).astype(np.int32))


# ----------------------------------------------------------------------------------------------------------
# the trace
# ----------------------------------------------------------------------------------------------------------
def trace(surfaces, materials, rays: np.ndarray, keep_all: bool = True, n_threads: int = 1) -> np.ndarray:
    """
    ``materials`` is the full list [initial] + system.materials + [final] (raytrace.py:653).
    rays (N, 8) -> (2S+1, N, 8) if keep_all else (N, 8).
    """
    if len(materials) != len(surfaces) + 1:
        raise ValueError("length of materials should be len(surfaces) + 1")
    rays = np.ascontiguousarray(rays, dtype=float)
    n = rays.shape[0]
    surf = pack_surfaces(surfaces)
    ntab, rows = index_table(materials, rays[:, 7])
    s = len(surfaces)
    out = np.empty((2 * s + 1, n, 8) if keep_all else (n, 8))
    lib().oracle_trace(_dptr(surf), s, _dptr(ntab), rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                       _dptr(rays), n, _dptr(out), int(keep_all), int(n_threads))
    return out


def prepared_trace(surfaces, materials, rays: np.ndarray, keep_all: bool = False, n_threads: int = 1):
    """
    For timing: do all Python-side packing once and return ``run() -> out`` that only executes the C loop
    (bench.py's cpu_baseline / --impl reference legs time ``run``).
    """
    rays = np.ascontiguousarray(rays, dtype=float)
    n = rays.shape[0]
    surf = pack_surfaces(surfaces)
    ntab, rows = index_table(materials, rays[:, 7])
    s = len(surfaces)
    out = np.zeros((2 * s + 1, n, 8) if keep_all else (n, 8))
    L = lib()
    rows_p = rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    def run():
        L.oracle_trace(_dptr(surf), s, _dptr(ntab), rows_p, _dptr(rays), n, _dptr(out), int(keep_all), int(n_threads))
        return out

    return run


def ray_trace(system, rays, initial_material, final_material, n_threads: int = 1) -> np.ndarray:
    """Same contract as System.ray_trace (raytrace.py:641-661), including (8,), (N,8) and (K,N,8) inputs."""
    materials = [initial_material] + list(system.materials) + [final_material]
    rays = np.asarray(rays, dtype=float)
    if rays.ndim == 1:
        rays = rays[None, None, :]
    elif rays.ndim == 2:
        rays = rays[None]
    hist = trace(system.surfaces, materials, rays[-1], keep_all=True, n_threads=n_threads)
    return np.concatenate((rays[:-1], hist), axis=0)


def sellmeier(b, c, wavelengths) -> np.ndarray:
    wl = np.ascontiguousarray(wavelengths, dtype=float).reshape(-1)
    out = np.empty_like(wl)
    bb = np.ascontiguousarray(b, dtype=float)
    cc = np.ascontiguousarray(c, dtype=float)
    lib().oracle_sellmeier(_dptr(bb), _dptr(cc), _dptr(wl), wl.size, _dptr(out))
    return out


def intersect_rays(ray1, ray2) -> np.ndarray:
    """intersect_rays, raytrace.py:164-238 (broadcast rules at 175-185)."""
    ray1 = np.atleast_2d(np.asarray(ray1, dtype=float))
    ray2 = np.atleast_2d(np.asarray(ray2, dtype=float))
    if len(ray1) == 1 and len(ray2) > 1:
        ray1 = np.tile(ray1, (len(ray2), 1))
    if len(ray2) == 1 and len(ray1) > 1:
        ray2 = np.tile(ray2, (len(ray1), 1))
    if len(ray1) != len(ray2):
        raise ValueError("ray1 and ray2 must be the same length")
    ray1 = np.ascontiguousarray(ray1)
    ray2 = np.ascontiguousarray(ray2)
    out = np.empty((len(ray1), 3))
    lib().oracle_intersect_rays(_dptr(ray1), _dptr(ray2), len(ray1), _dptr(out))
    return out


# ----------------------------------------------------------------------------------------------------------
# ray sources (NumPy restatements: the reference defines these through np.cos / np.sin / linspace)
# ----------------------------------------------------------------------------------------------------------
def fan_basis(center_ray):
    """raytrace.py:79-81"""
    center_ray = np.array(center_ray, dtype=float)
    e1 = np.cross(np.array([0, 1, 0]), center_ray)
    e1 = e1 / np.linalg.norm(e1)
    e2 = np.cross(center_ray, e1)
    return e1, e2


def collimated_basis(normal):
    """raytrace.py:135-144"""
    normal = np.array(normal, dtype=float).squeeze()
    n1 = np.cross(np.array([0, 1, 0]), normal)
    if np.linalg.norm(n1) == 0:
        n1 = np.cross(normal, np.array([1, 0, 0]))
    n1 = n1 / np.linalg.norm(n1)
    n2 = np.cross(normal, n1)
    n2 = n2 / np.linalg.norm(n2)
    return n1, n2


def source_rays(kind: str, n_a: int, n_b: int, a_max: float, pt, axis, wavelength: float, b_max: float = 0.0,
                b_start: float = 0.0, e1=None, e2=None, first: int = 0, count: int | None = None) -> np.ndarray:
    """
    Rows [first, first+count) of the index space of a device ray source (include/rtb.h rtb_source), computed the
    way the reference computes the corresponding generator:
      "fan"        get_ray_fan         raytrace.py:45-96    index = i_phi * n_a + i_theta
      "collimated" get_collimated_rays raytrace.py:99-161   index = i_disp * n_b + i_phi
      "grid"       this project's Cartesian pupil grid      index = i_v * n_a + i_u
    """
    total = n_a * n_b
    if count is None:
        count = total - first
    idx = np.arange(first, first + count, dtype=np.int64)
    pt = np.array(pt, dtype=float).squeeze()
    axis = np.array(axis, dtype=float).squeeze()
    rays = np.zeros((count, 8))
    a_vals = np.linspace(-a_max, a_max, n_a)
    if kind == "fan":
        if e1 is None:
            e1, e2 = fan_basis(axis)
        phis = np.arange(n_b) * 2 * np.pi / n_b
        tts = a_vals[idx % n_a]
        pps = phis[idx // n_a]
        rays[:, 0:3] = pt
        for k in range(3):
            rays[:, 3 + k] = (axis[k] * np.cos(tts) + e1[k] * np.cos(pps) * np.sin(tts) +
                              e2[k] * np.sin(pps) * np.sin(tts))
    elif kind == "collimated":
        if e1 is None:
            e1, e2 = collimated_basis(axis)
        phis = np.arange(n_b) * 2 * np.pi / n_b + b_start
        oos = a_vals[idx // n_b]
        pps = phis[idx % n_b]
        rays[:, 0:3] = (pt[None, :] + e1[None, :] * (oos * np.cos(pps))[:, None] +
                        e2[None, :] * (oos * np.sin(pps))[:, None])
        rays[:, 3:6] = axis
    elif kind == "grid":
        if e1 is None:
            e1, e2 = collimated_basis(axis)
        b_vals = np.linspace(-b_max, b_max, n_b)
        u = a_vals[idx % n_a]
        v = b_vals[idx // n_a]
        rays[:, 0:3] = (pt[None, :] + e1[None, :] * u[:, None]) + e2[None, :] * v[:, None]
        rays[:, 3:6] = axis
    else:
        raise ValueError(kind)
    rays[:, 6] = 0
    rays[:, 7] = wavelength
    return rays


# ----------------------------------------------------------------------------------------------------------
# this project's reductions (definitions in include/rtb.h, rtb_reduce)
# ----------------------------------------------------------------------------------------------------------
def reduce_stats(slab_rays: np.ndarray, origin, e1, e2, phase_ref: float = 0.0) -> np.ndarray:
    p = slab_rays[:, 0:3] - np.asarray(origin, dtype=float)
    u = p @ np.asarray(e1, dtype=float)
    v = p @ np.asarray(e2, dtype=float)
    ph = slab_rays[:, 6] - phase_ref
    ok = np.isfinite(u) & np.isfinite(v) & np.isfinite(ph)
    u, v, ph = u[ok], v[ok], ph[ok]
    if u.size == 0:
        return np.array([0, 0, 0, 0, 0, 0, 0, 0, np.inf, -np.inf, np.inf, -np.inf], dtype=float)
    return np.array([u.size, u.sum(), v.sum(), (u * u).sum(), (v * v).sum(), (u * v).sum(), ph.sum(),
                     (ph * ph).sum(), u.min(), u.max(), v.min(), v.max()], dtype=float)


def reduce_grid(slab_rays: np.ndarray, origin, e1, e2, grid_n: int, half_width: float,
                phase_ref: float = 0.0) -> np.ndarray:
    """(3, G, G): sum cos, sum sin, count; cell index = floor((u + half) * (G / (2*half))), [iv, iu] order."""
    p = slab_rays[:, 0:3] - np.asarray(origin, dtype=float)
    u = p @ np.asarray(e1, dtype=float)
    v = p @ np.asarray(e2, dtype=float)
    ph = slab_rays[:, 6] - phase_ref
    ok = np.isfinite(u) & np.isfinite(v) & np.isfinite(ph)
    inv_cell = grid_n / (2 * half_width)
    with np.errstate(invalid="ignore"):
        iu = np.floor((u + half_width) * inv_cell)
        iv = np.floor((v + half_width) * inv_cell)
        ok &= (iu >= 0) & (iu < grid_n) & (iv >= 0) & (iv < grid_n)
    iu = iu[ok].astype(np.int64)
    iv = iv[ok].astype(np.int64)
    ph = ph[ok]
    grid = np.zeros((3, grid_n, grid_n))
    np.add.at(grid[0], (iv, iu), np.cos(ph))
    np.add.at(grid[1], (iv, iu), np.sin(ph))
    np.add.at(grid[2], (iv, iu), 1.0)
    return grid
