"""
Host side of the trace: turns surface / material objects into the POD records of include/rtb.h and calls the C ABI.

Nothing in this module does ray arithmetic.  The only numerical work on the host is what the reference also does
per *system* rather than per ray: evaluating ``material.n()`` on the batch's distinct wavelengths (a handful of
numbers) and the per-surface constants ``radius**2``, ``abs(radius)``, ``normal*focal_len``, ``sin(alpha)``,
computed with the very Python expressions the reference uses so the kernel starts from identical bits
(reference raytrace.py:1499, 1528, 1683, 1758).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from .materials import KIND_TABLE_ONLY

SLAB_LAST = -1


# ----------------------------------------------------------------------------------------------------------
# packing
# ----------------------------------------------------------------------------------------------------------
def _v3(x):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size != 3:
        raise ValueError(f"expected a 3-vector, got shape {np.shape(x)}")
    return (C.c_double * 3)(float(a[0]), float(a[1]), float(a[2]))


def pack_surface(s) -> _ffi.RtbSurface:
    rec = getattr(s, "device_record", None)
    if rec is None:
        raise NotImplementedError(
            f"{type(s).__name__} has no device_record(): only FlatSurface, SphericalSurface, PlaneMirror and "
            f"PerfectLens (and subclasses that keep their geometry) can be traced; there is no CPU fallback")
    # a subclass that re-defines the ray arithmetic cannot be honoured by the kernel: refuse rather than ignore it
    from . import raytrace as _rt
    known = next((k for k in type(s).__mro__ if k in (_rt.FlatSurface, _rt.SphericalSurface, _rt.PlaneMirror,
                                                       _rt.PerfectLens)), None)
    if known is not None and type(s) is not known:
        for name in ("propagate", "get_intersect", "get_normal", "is_pt_on_surface"):
            if getattr(type(s), name) is not getattr(known, name):
                raise NotImplementedError(
                    f"{type(s).__name__} overrides {name}(); the GPU trace only implements {known.__name__}'s own "
                    f"geometry (no CPU fallback)")
    d = rec()
    out = _ffi.RtbSurface()
    out.kind = d["kind"]
    out.center = _v3(d["center"])
    out.normal = _v3(d["normal"])
    out.input_axis = _v3(d["input_axis"])
    out.radius = d.get("radius", 0.0)
    out.radius_sq = d.get("radius_sq", 0.0)
    out.abs_radius = d.get("abs_radius", 0.0)
    out.aperture_rad = d["aperture_rad"]
    out.focal_len = d.get("focal_len", 0.0)
    out.normal_f = _v3(d.get("normal_f", (0.0, 0.0, 0.0)))
    out.sin_alpha = d.get("sin_alpha", 0.0)
    return out


# ----------------------------------------------------------------------------------------------------------
# hints (rtb_surface.hints): never change a result, only which code path reaches it first
# ----------------------------------------------------------------------------------------------------------
def _first_flat(surfaces):
    from . import raytrace as _rt
    if not surfaces or not isinstance(surfaces[0], _rt.FlatSurface):
        return None
    s = surfaces[0]
    return np.asarray(s.center, dtype=np.float64).reshape(3), np.asarray(s.normal, dtype=np.float64).reshape(3)


def degenerate_first_surface(surfaces, rays=None, origin=None, direction=None) -> bool:
    """
    Does the bundle meet a FLAT first surface with exact zeros, ray after ray -- starting on the plane (t = +-0) or
    running exactly along its normal (d x n = 0)?  That is how the reference's scripts launch their rays (a flat at
    z = 0 under a collimated beam or a point source), and the kernel handles it in line when told
    (RTB_HINT_DEGENERATE) instead of redoing every ray out of line.

    ``rays``: a host sample of the batch, (n, 8); or ``origin`` / ``direction``: the point / direction every ray of an
    on-device source shares (either may be None).  A wrong answer only costs speed.
    """
    flat = _first_flat(surfaces)
    if flat is None:
        return False
    c, n = flat
    with np.errstate(all="ignore"):
        if rays is not None:
            r = np.asarray(rays, dtype=np.float64).reshape(-1, 8)
            r = r[np.isfinite(r[:, :6]).all(axis=1)]
            if r.shape[0] == 0:
                return False
            num = ((r[:, 0] - c[0]) * n[0] + (r[:, 1] - c[1]) * n[1]) + (r[:, 2] - c[2]) * n[2]
            d = r[:, 3:6]
            cross = np.stack((d[:, 1] * n[2] - d[:, 2] * n[1], d[:, 2] * n[0] - d[:, 0] * n[2],
                              d[:, 0] * n[1] - d[:, 1] * n[0]), axis=1)
            return bool((num == 0).all() or (cross == 0).all())
        on_plane = along = False
        if origin is not None:
            o = np.asarray(origin, dtype=np.float64).reshape(3)
            on_plane = bool(((o[0] - c[0]) * n[0] + (o[1] - c[1]) * n[1]) + (o[2] - c[2]) * n[2] == 0)
        if direction is not None:
            d = np.asarray(direction, dtype=np.float64).reshape(3)
            along = bool(d[1] * n[2] - d[2] * n[1] == 0 and d[2] * n[0] - d[0] * n[2] == 0
                         and d[0] * n[1] - d[1] * n[0] == 0)
        return on_plane or along


def set_first_surface_hint(packed, degenerate: bool):
    """mark / unmark surface 0 of a packed system (a PackedSystem is reused across launches by device.prepare)"""
    if packed.n_surfaces > 0:
        packed.sys.surfaces[0].hints = _ffi.HINT_DEGENERATE if degenerate else 0
    return packed


def pack_material(m) -> _ffi.RtbMaterial:
    rec = getattr(m, "device_record", None)
    out = _ffi.RtbMaterial()
    if rec is None:
        # a foreign object that only offers n(): host table only
        out.kind = KIND_TABLE_ONLY
        out.n_const = float("nan")
        return out
    kind, b, c, n_const = rec()
    out.kind = kind
    out.b = (C.c_double * 3)(*b)
    out.c = (C.c_double * 3)(*c)
    out.n_const = n_const
    return out


def distinct_wavelengths(wl: np.ndarray, limit: int):
    """
    Distinct non-NaN bit patterns of ``wl`` in order of first appearance, or None if there are more than ``limit``.
    O(N * distinct) with vectorised passes: typical batches carry 1-3 wavelengths.
    """
    bits = np.ascontiguousarray(wl, dtype=np.float64).view(np.int64)
    bits = bits[~np.isnan(wl)]
    found = []
    while bits.size:
        if len(found) == limit:
            return None
        v = bits[0]
        found.append(v)
        bits = bits[bits != v]
    return np.array(found, dtype=np.int64).view(np.float64)


def index_table(materials, wavelengths: np.ndarray) -> np.ndarray:
    """
    (U + 1, M) table: row u = [m.n(wavelengths[u]) for m in materials], last row = the materials' answer to a NaN
    wavelength.  Each material's own ``n`` is called once on the whole vector, as the reference calls it on arrays.
    """
    query = np.concatenate((np.asarray(wavelengths, dtype=np.float64).reshape(-1), [np.nan]))
    table = np.empty((query.size, len(materials)), dtype=np.float64)
    with np.errstate(all="ignore"):
        for j, m in enumerate(materials):
            try:
                col = np.asarray(m.n(query), dtype=np.float64).reshape(-1)
            except Exception:
                col = np.full(query.size, np.nan)
                head = np.asarray(m.n(query[:-1]), dtype=np.float64).reshape(-1)
                col[:-1] = np.broadcast_to(head, (query.size - 1,)) if head.size == 1 else head
            if col.size == 1:
                # a user material that answers an array with one number: the reference's arithmetic broadcasts it
                col = np.broadcast_to(col, (query.size,))
            if col.size != query.size:
                raise ValueError(f"{type(m).__name__}.n() returned {col.size} values for {query.size} wavelengths")
            table[:, j] = col
    return table


@dataclass
class PackedSystem:
    """RtbSystem plus the Python objects that own the memory it points to."""
    sys: _ffi.RtbSystem
    n_surfaces: int
    keep: list = field(default_factory=list)

    @property
    def n_slabs(self) -> int:
        return 2 * self.n_surfaces + 1


def pack_system(surfaces, materials, wavelengths=None) -> PackedSystem:
    """
    ``materials`` is the full list [initial] + system.materials + [final] (reference raytrace.py:653).
    ``wavelengths``: distinct wavelengths of the batch (-> host refractive-index table) or None (-> the kernel
    evaluates Sellmeier / constant media per ray).
    """
    S = len(surfaces)
    if len(materials) != S + 1:
        raise ValueError("length of materials should be len(surfaces) + 1")
    if S > _ffi.RTB_MAX_SURFACES:
        raise ValueError(f"at most {_ffi.RTB_MAX_SURFACES} surfaces per launch, got {S} (trace_host / trace_tensor split "
                         f"longer systems into segments)")
    surf_arr = (_ffi.RtbSurface * max(S, 1))(*[pack_surface(s) for s in surfaces])
    mat_arr = (_ffi.RtbMaterial * (S + 1))(*[pack_material(m) for m in materials])
    sys = _ffi.RtbSystem()
    sys.n_surfaces = S
    sys.surfaces = C.cast(surf_arr, C.POINTER(_ffi.RtbSurface))
    sys.materials = C.cast(mat_arr, C.POINTER(_ffi.RtbMaterial))
    keep = [surf_arr, mat_arr]
    if wavelengths is not None:
        wavelengths = np.ascontiguousarray(wavelengths, dtype=np.float64).reshape(-1)
        if wavelengths.size > _ffi.RTB_MAX_WAVELENGTHS:
            raise ValueError(f"at most {_ffi.RTB_MAX_WAVELENGTHS} tabulated wavelengths")
    if wavelengths is not None and wavelengths.size > 0:
        table = np.ascontiguousarray(index_table(materials, wavelengths))
        sys.n_wavelengths = wavelengths.size
        sys.wavelengths = wavelengths.ctypes.data_as(C.POINTER(C.c_double))
        sys.n_table = table.ctypes.data_as(C.POINTER(C.c_double))
        keep += [wavelengths, table]
    else:
        sys.n_wavelengths = 0
        bad = [type(m).__name__ for m, r in zip(materials, mat_arr) if r.kind == KIND_TABLE_ONLY]
        if bad:
            raise NotImplementedError(
                f"media {sorted(set(bad))} define their own n(); they are evaluated on the host per distinct wavelength, "
                f"which needs at most {_ffi.RTB_MAX_WAVELENGTHS} distinct wavelengths per batch (no CPU fallback)")
    return PackedSystem(sys, S, keep)


# Packing costs ~50 us per surface in Python, which is what a small launch is made of (1e6 rays x 3 surfaces trace in 45
# us).  Systems that are traced again and again -- sweeps, auto-focus iterations, benchmark loops -- come out of a small
# memo keyed by VALUE: the surface and material objects stay mutable, as in the reference, and a changed attribute is a
# different key.  Media that exist only as Python code (their table is whatever n() returns today) are never memoized.
_PACK_MEMO: "dict[tuple, PackedSystem]" = {}
_PACK_MEMO_SIZE = 32
_VECTOR_FIELDS = ("center", "normal", "input_axis", "normal_f")


def _surface_key(s):
    rec = getattr(s, "device_record", None)
    if rec is None:
        return None
    d = rec()
    key = [type(s)]
    for name, value in d.items():
        key.append(np.asarray(value, dtype=np.float64).tobytes() if name in _VECTOR_FIELDS else value)
    return tuple(key)


def _material_key(m):
    rec = getattr(m, "device_record", None)
    if rec is None:
        return None
    kind, b, c, n_const = rec()
    if kind == KIND_TABLE_ONLY:
        return None
    # (NaN never equals itself: keyed by bit pattern)
    return (type(m), kind, tuple(b), tuple(c), np.float64(n_const).tobytes())


def pack_system_memo(surfaces, materials, wavelengths=None) -> PackedSystem:
    """pack_system, memoized by the values it would pack (see above).  The returned object is shared between calls:
    callers may set surface hints on it before each launch, nothing else."""
    try:
        key = (tuple(_surface_key(s) for s in surfaces), tuple(_material_key(m) for m in materials),
               None if wavelengths is None else np.ascontiguousarray(wavelengths, dtype=np.float64).tobytes())
        if any(k is None for k in key[0]) or any(k is None for k in key[1]):
            key = None
        else:
            hash(key)
    except Exception:
        key = None          # anything unusual goes the plain way (and raises there what it has to raise)
    if key is None:
        return pack_system(surfaces, materials, wavelengths)
    hit = _PACK_MEMO.pop(key, None)
    if hit is None:
        hit = pack_system(surfaces, materials, wavelengths)
        while len(_PACK_MEMO) >= _PACK_MEMO_SIZE:
            _PACK_MEMO.pop(next(iter(_PACK_MEMO)))
    _PACK_MEMO[key] = hit       # (re-inserted: most recently used last)
    return hit


def resolve_keep(keep, n_slabs: int):
    """-> (keep_mode, int32 array or None, n_out_slabs)"""
    if isinstance(keep, str):
        if keep == "all":
            return _ffi.KEEP_ALL, None, n_slabs
        if keep == "last":
            return _ffi.KEEP_LAST, None, 1
        if keep == "none":
            return _ffi.KEEP_NONE, None, 0
        raise ValueError(f"keep must be 'all', 'last', 'none' or a list of slab indices, got {keep!r}")
    idx = [int(k) for k in keep]
    idx = [k + n_slabs if k < 0 else k for k in idx]
    if any(k < 0 or k >= n_slabs for k in idx):
        raise ValueError(f"slab indices must lie in [-{n_slabs}, {n_slabs})")
    if any(b <= a for a, b in zip(idx, idx[1:])):
        raise ValueError("slab indices must be strictly increasing")
    if not idx:
        return _ffi.KEEP_NONE, None, 0
    return _ffi.KEEP_LIST, np.array(idx, dtype=np.int32), len(idx)


def make_opts(keep_mode, keep_idx, precision="f64", reduce=None):
    opts = _ffi.RtbTraceOpts()
    try:
        opts.precision = {"f64": _ffi.F64_EXACT, "f32": _ffi.F32_FAST, "f64_fast": _ffi.F64_FAST}[precision]
    except KeyError:
        raise ValueError(f"precision must be 'f64', 'f64_fast' or 'f32', got {precision!r}") from None
    opts.keep_mode = keep_mode
    if keep_idx is not None:
        opts.n_keep = len(keep_idx)
        opts.keep_slabs = keep_idx.ctypes.data_as(C.POINTER(C.c_int32))
    if reduce is not None:
        opts.reduce = C.pointer(reduce)
    return opts


# ----------------------------------------------------------------------------------------------------------
# host-buffer trace (the drop-in call)
# ----------------------------------------------------------------------------------------------------------
def distinct_wavelengths_scan(rays: np.ndarray):
    """Complete list of distinct non-NaN wavelengths of a contiguous (N, 8) host batch (threaded C scan), or None
    when there are more than RTB_MAX_WAVELENGTHS."""
    out = np.empty(_ffi.RTB_MAX_WAVELENGTHS + 1, dtype=np.float64)
    found = C.c_int32(0)
    _ffi.check(_ffi.lib().rtb_distinct_wavelengths_host(rays.ctypes.data, rays.shape[0],
                                                        out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(found)))
    if found.value > _ffi.RTB_MAX_WAVELENGTHS:
        return None
    return out[:found.value].copy()


def choose_wavelength_table(materials, rays: np.ndarray):
    """
    Wavelengths to tabulate on the host for a contiguous (N, 8) host batch, or None for "evaluate in the kernel".

    * A medium that only exists as Python code (``Ebaf11``, user subclasses) needs the COMPLETE list: full scan.
    * Otherwise the kernel can evaluate any wavelength the table misses, so a 64-ray sample is enough to catch the
      usual "a few spectral lines per batch" case without reading the whole column on the host.
    """
    n = rays.shape[0]
    table_only = any(pack_material(m).kind == KIND_TABLE_ONLY for m in materials)
    if table_only:
        uniq = distinct_wavelengths_scan(rays)
        if uniq is None:
            return None          # pack_system raises the NotImplementedError naming the media
        return uniq if uniq.size else np.array([1.0])   # no valid wavelength at all: only the NaN row is used
    if n == 0:
        return None
    sample = rays[np.linspace(0, n - 1, num=min(n, 64)).astype(np.int64), 7]
    uniq = distinct_wavelengths(sample, _ffi.RTB_MAX_WAVELENGTHS)
    if uniq is None:
        return None
    return uniq if uniq.size else None


# ----------------------------------------------------------------------------------------------------------
# systems longer than one launch
# ----------------------------------------------------------------------------------------------------------
def global_keep_list(keep, n_slabs: int):
    """-> (sorted list of the trace's slab indices that are returned, n_out)"""
    mode, idx, n_out = resolve_keep(keep, n_slabs)
    if mode == _ffi.KEEP_ALL:
        return list(range(n_slabs)), n_out
    if mode == _ffi.KEEP_LAST:
        return [n_slabs - 1], n_out
    if mode == _ffi.KEEP_NONE:
        return [], n_out
    return [int(k) for k in idx], n_out


def plan_segments(n_surfaces: int, kept_slabs, reduce_slab=None, width: int | None = None):
    """
    A kernel launch carries at most RTB_MAX_SURFACES surfaces in its parameter block; the reference has no such limit
    (its loop, raytrace.py:657-659, just keeps appending slabs).  Longer systems are traced in segments: segment
    [a, b) starts from the last slab of the previous one -- exactly what the reference's loop hands to the next
    surface -- and contributes the trace's slabs 2a+1 .. 2b (the first segment also slab 0).

    Returns a list of ``(a, b, local, chain, local_reduce_slab)``: ``local`` = [(slab index inside the segment,
    position in the output or -1)], in increasing order; ``chain`` says the segment's last slab feeds the next one
    (it is then always the last entry of ``local``); ``local_reduce_slab`` is the fused reduction's slab inside
    this segment or None.
    """
    width = _ffi.RTB_MAX_SURFACES if width is None else int(width)
    position = {int(g): j for j, g in enumerate(kept_slabs)}
    plan = []
    for a in range(0, n_surfaces, width):
        b = min(n_surfaces, a + width)
        lo, hi = 2 * a, 2 * b
        local = [(g - lo, position[g]) for g in range(lo if a == 0 else lo + 1, hi + 1) if g in position]
        chain = b < n_surfaces
        if chain and (not local or local[-1][0] != hi - lo):
            local.append((hi - lo, -1))
        red = None
        if reduce_slab is not None and ((a == 0 and reduce_slab == 0) or lo < reduce_slab <= hi):
            red = reduce_slab - lo
        plan.append((a, b, local, chain, red))
    return plan


def reduce_for_segment(reduce, local_slab):
    """a copy of an RtbReduce whose slab index is relative to one segment"""
    one = _ffi.RtbReduce()
    C.memmove(C.byref(one), C.byref(reduce), C.sizeof(one))
    one.slab = int(local_slab)
    return one


def _trace_host_long(surfaces, materials, rays, keep, precision, device, reduce, out):
    S = len(surfaces)
    if len(materials) != S + 1:
        raise ValueError("length of materials should be len(surfaces) + 1")
    n_slabs = 2 * S + 1
    kept, n_out = global_keep_list(keep, n_slabs)
    n = rays.shape[0]
    if out is None:
        out = _ffi.result_array((n_out, n, 8))
    elif out.shape != (n_out, n, 8) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float64 array of shape {(n_out, n, 8)}")
    red_slab = None
    if reduce is not None:
        red_slab = int(reduce.slab) + (n_slabs if reduce.slab < 0 else 0)
    cur = rays
    for a, b, local, chain, red in plan_segments(S, kept, red_slab):
        if not local and red is None:
            continue
        part = trace_host(surfaces[a:b], materials[a:b + 1], cur, keep=[l for l, _ in local] or "none",
                          precision=precision, device=device,
                          reduce=None if red is None else reduce_for_segment(reduce, red))
        for j, (_, pos) in enumerate(local):
            if pos >= 0:
                out[pos] = part[j]
        if chain:
            cur = part[len(local) - 1]
    return out


def trace_host(surfaces, materials, rays: np.ndarray, keep="all", precision="f64", device: int = 0,
               reduce=None, out: np.ndarray | None = None) -> np.ndarray:
    """
    rays: (N, 8) float64 host array -> (n_out_slabs, N, 8) float64 host array.
    ``materials`` = [initial] + system.materials + [final].
    """
    _ffi.require_device()
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    if rays.ndim != 2 or rays.shape[1] != 8:
        raise ValueError(f"rays must have shape (N, 8), got {rays.shape}")
    n = rays.shape[0]
    if len(surfaces) > _ffi.RTB_MAX_SURFACES:
        return _trace_host_long(surfaces, materials, rays, keep, precision, device, reduce, out)
    uniq = choose_wavelength_table(materials, rays)
    if uniq is None and any(pack_material(m).kind == KIND_TABLE_ONLY for m in materials):
        return _trace_host_grouped(surfaces, materials, rays, keep, precision, device, reduce, out)
    packed = pack_system_memo(surfaces, materials, uniq)
    if n:
        sample = rays[np.linspace(0, n - 1, num=min(n, 64)).astype(np.int64)]
        set_first_surface_hint(packed, degenerate_first_surface(surfaces, rays=sample))
    mode, idx, n_out = resolve_keep(keep, packed.n_slabs)
    opts = make_opts(mode, idx, precision, reduce)
    if out is None:
        out = _ffi.result_array((n_out, n, 8))
    elif out.shape != (n_out, n, 8) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float64 array of shape {(n_out, n, 8)}")
    rc = _ffi.lib().rtb_trace_host(C.byref(packed.sys), rays.ctypes.data, n, out.ctypes.data if n_out else None,
                                   C.byref(opts), device)
    _ffi.check(rc)
    return out


MAX_WAVELENGTH_GROUPS = 16


def _trace_host_grouped(surfaces, materials, rays, keep, precision, device, reduce, out):
    """
    A batch with more than RTB_MAX_WAVELENGTHS distinct wavelengths AND media that only exist as Python code: the
    rays are traced in groups of RTB_MAX_WAVELENGTHS wavelengths, each group with its own complete host table
    (the GPU still does all the ray arithmetic; the host only selects and scatters rows).
    """
    wl = rays[:, 7]
    valid = ~np.isnan(wl)
    # grouped by BIT PATTERN, like the device table and the scan that sent the batch here (+0.0 and -0.0 are two
    # wavelengths to them; grouping by value would hand the same nine patterns back to trace_host for ever)
    wl_bits = np.ascontiguousarray(wl).view(np.int64)
    all_bits = np.unique(wl_bits[valid])
    all_wl = all_bits.view(np.float64)
    width = _ffi.RTB_MAX_WAVELENGTHS
    if all_wl.size > width * MAX_WAVELENGTH_GROUPS:
        bad = sorted({type(m).__name__ for m in materials if pack_material(m).kind == KIND_TABLE_ONLY})
        raise NotImplementedError(
            f"media {bad} define their own n() and the batch has {all_wl.size} distinct wavelengths; at most "
            f"{width * MAX_WAVELENGTH_GROUPS} are supported for host-evaluated media (no CPU fallback)")
    n = rays.shape[0]
    mode, idx, n_out = resolve_keep(keep, 2 * len(surfaces) + 1)
    if out is None:
        out = np.empty((n_out, n, 8), dtype=np.float64)
    elif out.shape != (n_out, n, 8):
        raise ValueError(f"out must have shape {(n_out, n, 8)}")
    for g in range(0, max(all_wl.size, 1), width):
        sel = valid & np.isin(wl_bits, all_bits[g:g + width])
        if g == 0:
            sel |= ~valid                        # NaN-wavelength rays ride with the first group (table's NaN row)
        if not sel.any():
            continue
        part = trace_host(surfaces, materials, np.ascontiguousarray(rays[sel]), keep=keep, precision=precision,
                          device=device, reduce=reduce)
        if n_out:
            out[:, sel, :] = part
    return out
