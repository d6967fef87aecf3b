"""
Device-resident API: the same kernels as ``System.ray_trace`` but with rays, histories and reduction buffers living
in HBM.  PyTorch is used only as the owner of device memory and streams (``torch.empty(..., device="cuda")``,
``tensor.data_ptr()``, ``torch.cuda.current_stream()``); every computation is a kernel of librtb.so.

Main pieces
-----------
``trace_tensor``      (N, 8) CUDA tensor -> (n_slabs, N, 8) CUDA tensor
``RaySource``         on-device ``get_ray_fan`` / ``get_collimated_rays`` / Cartesian grid (reference raytrace.py:45-161)
``trace_source``      source -> trace fused in one kernel: no input bytes at all
``Reducer``           spot statistics + pupil-grid accumulation fused into the trace (rtb_reduce in include/rtb.h)
``intersect_rays``, ``propagate_ray2plane``, ``surface_intersect``  the helpers around the trace
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi, engine


def _torch():
    import torch
    return torch


def _stream_ptr(device_index: int) -> int:
    return _torch().cuda.current_stream(device_index).cuda_stream


def _check_rays_tensor(t, what="rays", planes=False):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{what} must be a CUDA torch.Tensor")
    if t.dtype != torch.float64:
        raise TypeError(f"{what} must be float64, got {t.dtype}")
    if planes:
        if t.dim() != 2 or t.shape[0] != 8:
            raise ValueError(f"{what} must have shape (8, N) in the plane layout, got {tuple(t.shape)}")
    elif t.dim() != 2 or t.shape[1] != 8:
        raise ValueError(f"{what} must have shape (N, 8), got {tuple(t.shape)}")
    return t.contiguous()


def distinct_wavelengths_tensor(rays):
    """Distinct non-NaN wavelengths of a device batch (ascending), or None if more than RTB_MAX_WAVELENGTHS."""
    torch = _torch()
    rays = _check_rays_tensor(rays)
    dev = rays.device.index
    scratch = torch.empty(_ffi.RTB_MAX_WAVELENGTHS + 1, dtype=torch.float64, device=rays.device)
    host = np.empty(_ffi.RTB_MAX_WAVELENGTHS + 1, dtype=np.float64)
    found = C.c_int32(0)
    rc = _ffi.lib().rtb_distinct_wavelengths_device(rays.data_ptr(), rays.shape[0], scratch.data_ptr(),
                                                    host.ctypes.data_as(C.POINTER(C.c_double)), C.byref(found), dev,
                                                    _stream_ptr(dev))
    _ffi.check(rc)
    if found.value > _ffi.RTB_MAX_WAVELENGTHS:
        return None
    return host[:found.value].copy()


def _system_for(surfaces, materials, wavelengths):
    """wavelengths: array of distinct values, None -> in-kernel Sellmeier."""
    if wavelengths is not None:
        wavelengths = np.asarray(wavelengths, dtype=np.float64).reshape(-1)
        if wavelengths.size == 0:
            wavelengths = None
    if wavelengths is None and any(engine.pack_material(m).kind == _ffi.MAT_TABLE_ONLY for m in materials):
        wavelengths = np.array([1.0])   # batch without a valid wavelength: only the NaN row is ever used
    return engine.pack_system_memo(surfaces, materials, wavelengths)


def trace_tensor(surfaces, materials, rays, keep="all", precision="f64", wavelengths="auto", reducer=None,
                 out=None, flags: int = 0, layout: str = "rows", degenerate_first=None):
    """
    rays: (N, 8) float64 CUDA tensor; ``materials`` = [initial] + system.materials + [final].
    wavelengths: "auto" scans the batch on the device for its distinct wavelengths (one extra pass over column 7);
    pass the known values (e.g. ``[0.785]``) to skip that, or None to force in-kernel Sellmeier evaluation.
    layout: "rows" = the reference's (N, 8) in, (n_slabs, N, 8) out; "planes" = structure-of-arrays, (8, N) in and
    (n_slabs, 8, N) out (one contiguous plane per column; give ``wavelengths`` explicitly or it is scanned from a
    row-major copy of the wavelength plane).
    degenerate_first: True / False = the bundle does / does not meet a flat first surface with exact zeros ray after
    ray (engine.degenerate_first_surface; a speed hint only), "auto" = decide from a 64-ray sample copied to the host
    (one small synchronising read), None = no hint -- except with wavelengths="auto", which synchronises anyway and
    therefore also samples for the hint.
    Returns a CUDA tensor on the same device (None for keep="none").  Enqueued on the current stream.
    """
    torch = _torch()
    if layout not in ("rows", "planes"):
        raise ValueError("layout must be 'rows' or 'planes'")
    planes = layout == "planes"
    rays = _check_rays_tensor(rays, planes=planes)
    dev = rays.device.index
    if isinstance(wavelengths, str):
        if wavelengths != "auto":
            raise ValueError("wavelengths must be 'auto', None or a sequence of values")
        if degenerate_first is None:
            degenerate_first = "auto"       # this route reads from the device anyway: the hint's sample rides along
        if planes:
            probe = torch.zeros((rays.shape[1], 8), dtype=torch.float64, device=rays.device)
            probe[:, 7] = rays[7]
            wavelengths = distinct_wavelengths_tensor(probe)
        else:
            wavelengths = distinct_wavelengths_tensor(rays)
    if isinstance(degenerate_first, str):
        if degenerate_first != "auto":
            raise ValueError("degenerate_first must be True, False, None or 'auto'")
        n_all = rays.shape[1] if planes else rays.shape[0]
        degenerate_first = False
        if n_all:
            pick = torch.linspace(0, n_all - 1, min(n_all, 64), device=rays.device).long()
            sample = (rays[:, pick].t() if planes else rays[pick]).cpu().numpy()
            degenerate_first = engine.degenerate_first_surface(surfaces, rays=sample)
    if len(surfaces) > _ffi.RTB_MAX_SURFACES:
        return _trace_tensor_long(surfaces, materials, rays, keep, precision, wavelengths, reducer, out, flags, layout,
                                  degenerate_first)
    packed = _system_for(surfaces, materials, wavelengths)
    engine.set_first_surface_hint(packed, bool(degenerate_first))
    mode, idx, n_out = engine.resolve_keep(keep, packed.n_slabs)
    opts = engine.make_opts(mode, idx, precision, reducer.struct if reducer is not None else None)
    opts.flags = flags | ((_ffi.FLAG_PLANES_IN | _ffi.FLAG_PLANES_OUT) if planes else 0)
    n = rays.shape[1] if planes else rays.shape[0]
    out_shape = (n_out, 8, n) if planes else (n_out, n, 8)
    if n_out == 0:
        out = None
    elif out is None:
        out = torch.empty(out_shape, dtype=torch.float64, device=rays.device)
    elif tuple(out.shape) != out_shape or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float64 tensor of shape {out_shape}")
    rc = _ffi.lib().rtb_trace_device(C.byref(packed.sys), rays.data_ptr(), n, out.data_ptr() if out is not None else None,
                                     C.byref(opts), dev, _stream_ptr(dev))
    _ffi.check(rc)
    return out


class _SegmentReducer:
    """a Reducer seen from one segment of a long system: same buffers, slab index relative to the segment"""

    def __init__(self, reducer, local_slab):
        self.struct = engine.reduce_for_segment(reducer.struct, local_slab)


def _trace_tensor_long(surfaces, materials, rays, keep, precision, wavelengths, reducer, out, flags, layout,
                       degenerate_first=None):
    """A system of more than RTB_MAX_SURFACES surfaces, in segments chained on the device (engine.plan_segments)."""
    torch = _torch()
    planes = layout == "planes"
    S = len(surfaces)
    if len(materials) != S + 1:
        raise ValueError("length of materials should be len(surfaces) + 1")
    n_slabs = 2 * S + 1
    kept, n_out = engine.global_keep_list(keep, n_slabs)
    n = rays.shape[1] if planes else rays.shape[0]
    out_shape = (n_out, 8, n) if planes else (n_out, n, 8)
    if n_out == 0:
        out = None
    elif out is None:
        out = torch.empty(out_shape, dtype=torch.float64, device=rays.device)
    elif tuple(out.shape) != out_shape or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float64 tensor of shape {out_shape}")
    red_slab = None
    if reducer is not None:
        red_slab = int(reducer.struct.slab) + (n_slabs if reducer.struct.slab < 0 else 0)
    cur = rays
    for a, b, local, chain, red in engine.plan_segments(S, kept, red_slab):
        if not local and red is None:
            continue
        part = trace_tensor(surfaces[a:b], materials[a:b + 1], cur, keep=[l for l, _ in local] or "none",
                            precision=precision, wavelengths=wavelengths,
                            reducer=None if red is None else _SegmentReducer(reducer, red), flags=flags, layout=layout,
                            degenerate_first=degenerate_first if a == 0 else None)
        for j, (_, pos) in enumerate(local):
            if pos >= 0:
                out[pos].copy_(part[j])
        if chain:
            cur = part[len(local) - 1]
    return out


# ----------------------------------------------------------------------------------------------------------
# ray sources
# ----------------------------------------------------------------------------------------------------------
@dataclass
class RaySource:
    """
    Description of a structured ray bundle that the device can enumerate by index (rtb_source in include/rtb.h).
    Build one with :meth:`fan`, :meth:`collimated` or :meth:`grid`; the transverse basis is computed here with the
    reference's own expressions (raytrace.py:79-81, 135-144).
    """
    kind: int
    n_a: int
    n_b: int
    a_max: float
    b_max: float
    b_start: float
    pt: np.ndarray
    axis: np.ndarray
    e1: np.ndarray
    e2: np.ndarray
    wavelength: float

    @property
    def n_rays(self) -> int:
        return self.n_a * self.n_b

    def degenerate_at_first(self, surfaces) -> bool:
        """engine.degenerate_first_surface for this source: a fan's rays share their origin, the rays of a collimated
        bundle or grid their direction (and lie in the plane through ``pt`` across it)."""
        if self.kind == _ffi.SRC_FAN:
            return engine.degenerate_first_surface(surfaces, origin=self.pt)
        return engine.degenerate_first_surface(surfaces, direction=self.axis)

    @classmethod
    def fan(cls, pt, theta_max, n_thetas, wavelength, nphis=1, center_ray=(0, 0, 1)):
        """get_ray_fan (raytrace.py:45-96): index = i_phi * n_thetas + i_theta"""
        from .raytrace import _transverse_basis
        axis = np.array(center_ray, dtype=float)
        if np.linalg.norm(axis) != 1:
            raise ValueError("center_ray must be a unit vector")
        e1, e2 = _transverse_basis(axis, fallback=False)
        return cls(_ffi.SRC_FAN, int(n_thetas), int(nphis), float(theta_max), 0.0, 0.0,
                   np.array(pt, dtype=float).reshape(3), axis, e1, e2, float(wavelength))

    @classmethod
    def collimated(cls, pt, displacement_max, n_disps, wavelength, nphis=1, phi_start=0., normal=(0, 0, 1)):
        """get_collimated_rays (raytrace.py:99-161): index = i_disp * nphis + i_phi"""
        from .raytrace import _transverse_basis
        if np.abs(np.linalg.norm(normal) - 1) > 1e-12:
            raise ValueError("normal must be a normalized vector")
        axis = np.array(normal, dtype=float).reshape(3)
        e1, e2 = _transverse_basis(axis, fallback=True)
        return cls(_ffi.SRC_COLLIMATED, int(n_disps), int(nphis), float(displacement_max), 0.0, float(phi_start),
                   np.array(pt, dtype=float).reshape(3), axis, e1, e2, float(wavelength))

    @classmethod
    def grid(cls, pt, half_width_u, n_u, wavelength, half_width_v=None, n_v=None, normal=(0, 0, 1)):
        """Cartesian grid of parallel rays, u along e1 and v along e2: index = i_v * n_u + i_u"""
        from .raytrace import _transverse_basis
        if np.abs(np.linalg.norm(normal) - 1) > 1e-12:
            raise ValueError("normal must be a normalized vector")
        axis = np.array(normal, dtype=float).reshape(3)
        e1, e2 = _transverse_basis(axis, fallback=True)
        return cls(_ffi.SRC_GRID, int(n_u), int(n_u if n_v is None else n_v), float(half_width_u),
                   float(half_width_u if half_width_v is None else half_width_v), 0.0,
                   np.array(pt, dtype=float).reshape(3), axis, e1, e2, float(wavelength))

    @property
    def struct(self) -> _ffi.RtbSource:
        s = _ffi.RtbSource()
        s.kind, s.n_a, s.n_b = self.kind, self.n_a, self.n_b
        s.a_max, s.b_max, s.b_start = self.a_max, self.b_max, self.b_start
        s.pt = engine._v3(self.pt)
        s.axis = engine._v3(self.axis)
        s.e1 = engine._v3(self.e1)
        s.e2 = engine._v3(self.e2)
        s.wavelength = self.wavelength
        return s

    def generate(self, first: int = 0, count: int | None = None, device: int = 0):
        """Materialise rays [first, first+count) as an (count, 8) CUDA tensor."""
        torch = _torch()
        _ffi.require_device()
        count = self.n_rays - first if count is None else count
        out = torch.empty((count, 8), dtype=torch.float64, device=f"cuda:{device}")
        src = self.struct
        rc = _ffi.lib().rtb_generate_device(C.byref(src), first, count, out.data_ptr(), device, _stream_ptr(device))
        _ffi.check(rc)
        return out


def prepare(surfaces, materials, wavelengths):
    """
    Pack a prescription once for many launches (sweeps over fields / chunks): returns an opaque object accepted as
    ``packed=`` by :func:`trace_source`.  ``wavelengths``: the distinct wavelengths to tabulate (or None).
    """
    return _system_for(surfaces, materials, wavelengths)


def trace_source(surfaces, materials, source: RaySource, first: int = 0, count: int | None = None, keep="last",
                 precision="f64", reducer=None, device: int = 0, out=None, packed=None):
    """
    Generate-and-trace in one kernel: rays [first, first+count) of ``source`` never exist in memory.
    The refractive-index table is built from the source's single wavelength (or taken from ``packed``, see
    :func:`prepare`, which must tabulate that wavelength).  Enqueued on the current stream of ``device``.
    """
    torch = _torch()
    _ffi.require_device()
    count = source.n_rays - first if count is None else count
    if packed is None and len(surfaces) > _ffi.RTB_MAX_SURFACES:
        # longer than one launch: the rays are generated into memory and the segments chained (trace_tensor)
        rays = source.generate(first, count, device=device)
        return trace_tensor(surfaces, materials, rays, keep=keep, precision=precision,
                            wavelengths=[source.wavelength] if np.isfinite(source.wavelength) else None,
                            reducer=reducer, out=out, degenerate_first=source.degenerate_at_first(surfaces))
    if packed is None:
        packed = _system_for(surfaces, materials, [source.wavelength] if np.isfinite(source.wavelength) else None)
    engine.set_first_surface_hint(packed, source.degenerate_at_first(surfaces))
    mode, idx, n_out = engine.resolve_keep(keep, packed.n_slabs)
    opts = engine.make_opts(mode, idx, precision, reducer.struct if reducer is not None else None)
    out = _check_out(out, (n_out, count, 8), device)
    src = source.struct
    rc = _ffi.lib().rtb_trace_source(C.byref(packed.sys), C.byref(src), first, count,
                                     out.data_ptr() if out is not None else None, C.byref(opts), device,
                                     _stream_ptr(device))
    _ffi.check(rc)
    return out


def _check_out(out, shape, device):
    """the caller's output tensor, validated (a wrong shape would be an out-of-bounds device write), or a fresh one"""
    torch = _torch()
    if shape[0] == 0:
        return None
    if out is None:
        return torch.empty(shape, dtype=torch.float64, device=f"cuda:{device}")
    if (tuple(out.shape) != tuple(shape) or out.dtype != torch.float64 or not out.is_contiguous() or not out.is_cuda
            or out.device.index != device):
        raise ValueError(f"out must be a contiguous float64 CUDA tensor of shape {tuple(shape)} on device {device}")
    return out


def _trace_sources_grouped(surfaces, materials, sources, wls, first, count, keep, precision, reducer, device, out):
    """A sweep over more than RTB_MAX_WAVELENGTHS wavelengths through media that need a host table: the sources are
    launched in groups of <= 8 wavelengths, each into its own rows of the output and its own reduction buckets."""
    torch = _torch()
    mode, idx, n_out = engine.resolve_keep(keep, 2 * len(surfaces) + 1)
    out = _check_out(out, (n_out, len(sources) * count, 8), device)
    width = _ffi.RTB_MAX_WAVELENGTHS
    for g in range(0, len(wls), width):
        group = set(wls[g:g + width])
        members = [k for k, s in enumerate(sources) if float(s.wavelength) in group
                   or (g == 0 and not np.isfinite(s.wavelength))]
        if not members:
            continue
        sub = Reducer.view_of(reducer, members) if reducer is not None else None
        part = trace_sources(surfaces, materials, [sources[k] for k in members], first=first, count=count, keep=keep,
                             precision=precision, reducer=sub, device=device,
                             packed=_system_for(surfaces, materials, sorted(group)))
        if sub is not None:
            sub.scatter_back()
        if out is not None:
            for j, k in enumerate(members):
                out[:, k * count:(k + 1) * count] = part[:, j * count:(j + 1) * count]
    return out


def trace_sources(surfaces, materials, sources, first: int = 0, count: int | None = None, keep="none",
                  precision="f64", reducer=None, device: int = 0, out=None, packed=None):
    """
    A sweep in ONE launch (kernel grid y = source): every ``RaySource`` of ``sources`` -- field points, wavelengths,
    defocused object points ... -- generates rays [first, first+count) of its own index space and traces them through
    the same system.  Source k fills rows [k*count, (k+1)*count) of every kept slab (``out`` is
    ``(n_kept, len(sources)*count, 8)``) and bucket k of ``reducer`` (create it with ``buckets=len(sources)``).
    The sources must have the same number of rays; their wavelengths are tabulated together (at most 8 distinct ones
    unless every medium has a closed formula).
    """
    torch = _torch()
    _ffi.require_device()
    sources = list(sources)
    if not sources:
        raise ValueError("no sources")
    if len({s.n_rays for s in sources}) != 1:
        raise ValueError("the sources of one sweep launch must have the same number of rays")
    count = sources[0].n_rays - first if count is None else count
    if reducer is not None and reducer.buckets != len(sources):
        raise ValueError(f"reducer has {reducer.buckets} bucket(s), the sweep has {len(sources)} sources")
    if packed is None:
        wls = sorted({float(s.wavelength) for s in sources if np.isfinite(s.wavelength)})
        if len(wls) > _ffi.RTB_MAX_WAVELENGTHS:
            if any(engine.pack_material(m).kind == engine.KIND_TABLE_ONLY for m in materials):
                # media that only exist as Python code need every wavelength tabulated: one launch per group of 8
                return _trace_sources_grouped(surfaces, materials, sources, wls, first, count, keep, precision, reducer,
                                              device, out)
            wls = None      # every medium has a closed formula: the kernel evaluates it per ray
        packed = _system_for(surfaces, materials, wls or None)
    engine.set_first_surface_hint(packed, all(s.degenerate_at_first(surfaces) for s in sources))
    mode, idx, n_out = engine.resolve_keep(keep, packed.n_slabs)
    opts = engine.make_opts(mode, idx, precision, reducer.struct if reducer is not None else None)
    out = _check_out(out, (n_out, len(sources) * count, 8), device)
    arr = (_ffi.RtbSource * len(sources))(*[s.struct for s in sources])
    rc = _ffi.lib().rtb_trace_sources(C.byref(packed.sys), arr, len(sources), first, count,
                                      out.data_ptr() if out is not None else None, C.byref(opts), device,
                                      _stream_ptr(device))
    _ffi.check(rc)
    return out


# ----------------------------------------------------------------------------------------------------------
# fused reductions
# ----------------------------------------------------------------------------------------------------------
class Reducer:
    """
    Spot statistics and pupil-grid accumulation sampled at one slab of the trace, fused into the trace kernel.

    u = (p - origin).e1, v = (p - origin).e2.  ``stats`` is the 12-vector documented in include/rtb.h; ``grid`` is a
    (3, G, G) tensor: sum cos(phase - phase_ref), sum sin(phase - phase_ref), count (indexed [plane, iv, iu]).
    Accumulates across calls until :meth:`reset`.  ``allreduce()`` sums over the ranks of the default
    ``torch.distributed`` group (NCCL over NVLink): the only communication of a multi-GPU trace.
    ``buckets`` > 1 makes one set per source of a sweep (:func:`trace_sources`): ``stats_t`` is (buckets, 12) and
    ``grid_t`` (buckets, 3, G, G).
    ``phase_ref`` is a global phase subtracted before cos / sin and the phase sums are formed.  Give it the chief ray's
    phase (``analysis.reference_phase``; ``analysis.pupil_grid`` does so by default): accumulated phases reach 1e7 rad
    and with ``phase_ref = 0`` every ray pays for a large-argument range reduction in the kernel.
    """

    def __init__(self, slab: int, origin=(0, 0, 0), e1=(1, 0, 0), e2=(0, 1, 0), grid_n: int = 0,
                 half_width: float = 1.0, phase_ref: float = 0.0, stats: bool = True, device: int = 0,
                 buckets: int = 1):
        torch = _torch()
        _ffi.require_device()
        if buckets < 1:
            raise ValueError("buckets must be >= 1")
        self.device = device
        self.slab = int(slab)
        self.grid_n = int(grid_n)
        self.buckets = int(buckets)
        lead = (self.buckets,) if self.buckets > 1 else ()
        self.stats_t = (torch.empty(lead + (_ffi.RTB_N_STATS,), dtype=torch.float64, device=f"cuda:{device}")
                        if stats else None)
        self.grid_t = (torch.empty(lead + (3, grid_n, grid_n), dtype=torch.float64, device=f"cuda:{device}")
                       if grid_n > 0 else None)
        s = _ffi.RtbReduce()
        s.slab = self.slab
        s.grid_n = self.grid_n
        s.origin = engine._v3(origin)
        s.e1 = engine._v3(e1)
        s.e2 = engine._v3(e2)
        s.phase_ref = float(phase_ref)
        s.grid_half_width = float(half_width)
        s.stats_dev = self.stats_t.data_ptr() if stats else None
        s.grid_dev = self.grid_t.data_ptr() if grid_n > 0 else None
        self.struct = s
        self.reset()

    @classmethod
    def view_of(cls, parent, members):
        """A reducer over buckets ``members`` of ``parent`` (its own contiguous storage, initialised with the parent's
        current contents); :meth:`scatter_back` writes the buckets back.  Lets a sweep be launched in groups."""
        st = parent.struct
        sub = cls(st.slab, origin=tuple(st.origin), e1=tuple(st.e1), e2=tuple(st.e2), grid_n=parent.grid_n,
                  half_width=st.grid_half_width, phase_ref=st.phase_ref, stats=parent.stats_t is not None,
                  device=parent.device, buckets=len(members))
        idx = _torch().as_tensor(list(members), device=f"cuda:{parent.device}")
        lead = (lambda t: t if parent.buckets > 1 else t.unsqueeze(0))
        if sub.stats_t is not None:
            sub.stats_t.reshape(len(members), -1).copy_(lead(parent.stats_t).index_select(0, idx))
        if sub.grid_t is not None:
            sub.grid_t.reshape((len(members),) + tuple(lead(parent.grid_t).shape[1:])).copy_(
                lead(parent.grid_t).index_select(0, idx))
        sub._parent, sub._members = parent, idx
        return sub

    def scatter_back(self):
        parent, idx = self._parent, self._members
        lead = (lambda t: t if parent.buckets > 1 else t.unsqueeze(0))
        if self.stats_t is not None:
            lead(parent.stats_t).index_copy_(0, idx, self.stats_t.reshape(len(idx), -1))
        if self.grid_t is not None:
            lead(parent.grid_t).index_copy_(0, idx, self.grid_t.reshape((len(idx),) + tuple(lead(parent.grid_t).shape[1:])))

    def resolve_slab(self, n_slabs: int):
        """allow negative slab indices once the system is known"""
        if self.struct.slab < 0:
            self.struct.slab += n_slabs
        return self

    def reset(self):
        for b in range(self.buckets):
            one = _ffi.RtbReduce()
            C.memmove(C.byref(one), C.byref(self.struct), C.sizeof(one))
            if self.stats_t is not None:
                one.stats_dev = self.stats_t.data_ptr() + b * _ffi.RTB_N_STATS * 8
            if self.grid_t is not None:
                one.grid_dev = self.grid_t.data_ptr() + b * 3 * self.grid_n * self.grid_n * 8
            _ffi.check(_ffi.lib().rtb_reduce_init(C.byref(one), self.device, _stream_ptr(self.device)))

    def allreduce(self, comm=None):
        """
        Sum the grid and merge the statistics over the ranks, in place, enqueued on the current stream.  ``comm``: a
        :class:`ray_trace_pb_b200.sharding.Comm` (the library's own NCCL communicator behind the C ABI: one all-reduce
        for the grid, one all-gather + merge kernel for the statistics); without it the default torch.distributed
        process group is used (any backend -- the CPU tests run this route over gloo).
        """
        if comm is not None:
            stream = _stream_ptr(self.device)
            if self.grid_t is not None:
                comm.allreduce_grid(self.grid_t, stream)
            if self.stats_t is not None:
                comm.allreduce_stats(self.stats_t, stream)
            return self
        from .sharding import allreduce_grid, allreduce_stats
        if self.grid_t is not None:
            allreduce_grid(self.grid_t)
        if self.stats_t is not None:
            allreduce_stats(self.stats_t)
        return self

    def psf(self, n_samples: int, df: float, normalize_by_count: bool = False, field: bool = False, bucket: int = 0):
        """
        PSF from the accumulated pupil grid as a zoomed DFT (kernel zgemm_nt_kernel): ``n_samples x n_samples`` samples
        of |E|^2, E = A P B^T, at spatial frequencies (k - (n-1)/2) * df along u (columns) and v (rows).  For a lens
        of focal length f the image coordinate is ``wavelength * f * frequency``.  Returns a CUDA tensor (n, n), or
        (psf, field_complex) with ``field=True``.
        """
        torch = _torch()
        if self.grid_t is None:
            raise ValueError("this Reducer was created without a grid")
        L = _ffi.lib()
        need = L.rtb_psf_scratch_doubles(self.grid_n, n_samples, int(normalize_by_count))
        dev = f"cuda:{self.device}"
        scratch = torch.empty(need, dtype=torch.float64, device=dev)
        out = torch.empty((n_samples, n_samples), dtype=torch.float64, device=dev)
        f_re = torch.empty_like(out) if field else None
        f_im = torch.empty_like(out) if field else None
        if not 0 <= bucket < self.buckets:
            raise IndexError(f"bucket {bucket} of {self.buckets}")
        grid_ptr = self.grid_t.data_ptr() + bucket * 3 * self.grid_n * self.grid_n * 8
        rc = L.rtb_psf_from_grid_device(grid_ptr, self.grid_n, float(self.struct.grid_half_width),
                                        n_samples, float(df), int(normalize_by_count), scratch.data_ptr(), need,
                                        out.data_ptr(), f_re.data_ptr() if field else None,
                                        f_im.data_ptr() if field else None, self.device, _stream_ptr(self.device))
        _ffi.check(rc)
        return (out, torch.complex(f_re, f_im)) if field else out

    def stats(self, bucket: int | None = None):
        """Host copy of the statistics with the derived centroid / RMS radius / RMS wavefront error (a list with one
        entry per bucket for a sweep reducer, unless ``bucket`` picks one)."""
        v = self.stats_t.cpu().numpy()
        if self.buckets == 1:
            return summarize_stats(v)
        if bucket is not None:
            return summarize_stats(v[bucket])
        return [summarize_stats(row) for row in v]

    @property
    def grid(self):
        return self.grid_t


def summarize_stats(v: np.ndarray) -> dict:
    n = v[0]
    out = {"count": int(n), "raw": v.copy()}
    if n > 0:
        mu, mv = v[1] / n, v[2] / n
        out["centroid"] = (mu, mv)
        out["rms_radius"] = float(np.sqrt(max(v[3] / n - mu * mu + v[4] / n - mv * mv, 0.0)))
        mp = v[6] / n
        out["mean_phase"] = mp
        out["rms_phase"] = float(np.sqrt(max(v[7] / n - mp * mp, 0.0)))
        out["u_range"] = (v[8], v[9])
        out["v_range"] = (v[10], v[11])
    return out


# ----------------------------------------------------------------------------------------------------------
# helpers around the trace
# ----------------------------------------------------------------------------------------------------------
def _to_device(a, device=0):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a.to(device=f"cuda:{device}", dtype=torch.float64).contiguous(), True
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if not arr.flags.writeable:           # e.g. a broadcast view; torch.from_numpy wants a writable buffer
        arr = arr.copy()
    return torch.from_numpy(arr).to(f"cuda:{device}"), False


def intersect_rays(ray1, ray2, device: int = 0):
    """intersect_rays (raytrace.py:164-238) on the device.  NumPy in -> NumPy (N, 3) out; CUDA tensors stay tensors."""
    torch = _torch()
    _ffi.require_device()
    as_np = not (isinstance(ray1, torch.Tensor) or isinstance(ray2, torch.Tensor))
    if as_np:
        ray1 = np.atleast_2d(np.asarray(ray1, dtype=np.float64))
        ray2 = np.atleast_2d(np.asarray(ray2, dtype=np.float64))
    else:
        ray1 = ray1 if ray1.dim() == 2 else ray1.reshape(1, -1)
        ray2 = ray2 if ray2.dim() == 2 else ray2.reshape(1, -1)
        if isinstance(ray1, torch.Tensor) and ray1.is_cuda:
            device = ray1.device.index
    n1, n2 = len(ray1), len(ray2)
    if not (n1 == n2 or n1 == 1 or n2 == 1):
        raise ValueError("ray1 and ray2 must be the same length")
    t1, _ = _to_device(ray1, device)
    t2, _ = _to_device(ray2, device)
    n = max(n1, n2)
    out = torch.empty((n, 3), dtype=torch.float64, device=f"cuda:{device}")
    rc = _ffi.lib().rtb_intersect_rays_device(t1.data_ptr(), n1, t2.data_ptr(), n2, out.data_ptr(), device,
                                              _stream_ptr(device))
    _ffi.check(rc)
    return out.cpu().numpy() if as_np else out


def propagate_ray2plane(rays, normal, center, material, exclude_backward_propagation: bool = False, device: int = 0):
    """propagate_ray2plane (raytrace.py:241-306): returns (rays_out (N, 8), ts (N,)) as NumPy arrays."""
    torch = _torch()
    _ffi.require_device()
    rays = np.atleast_2d(np.array(rays, dtype=np.float64, copy=True))
    n = rays.shape[0]
    normal = np.atleast_2d(np.array(normal, dtype=np.float64).squeeze())
    center = np.atleast_2d(np.array(center, dtype=np.float64).squeeze())
    for name, arr in (("normal", normal), ("center", center)):
        if arr.shape[-1] != 3 or arr.shape[0] not in (1, n):
            raise ValueError(f"{name} must be broadcastable to ({n}, 3)")
    with np.errstate(all="ignore"):
        index = np.broadcast_to(np.asarray(material.n(rays[:, 7]), dtype=np.float64).reshape(-1), (n,))
    t_rays, _ = _to_device(rays, device)
    t_n, _ = _to_device(normal, device)
    t_c, _ = _to_device(center, device)
    t_i, _ = _to_device(np.ascontiguousarray(index), device)
    out = torch.empty((n, 8), dtype=torch.float64, device=f"cuda:{device}")
    ts = torch.empty((n,), dtype=torch.float64, device=f"cuda:{device}")
    rc = _ffi.lib().rtb_ray2plane_device(t_rays.data_ptr(), n, t_n.data_ptr(), normal.shape[0], t_c.data_ptr(),
                                         center.shape[0], t_i.data_ptr(), int(bool(exclude_backward_propagation)),
                                         out.data_ptr(), ts.data_ptr(), device, _stream_ptr(device))
    _ffi.check(rc)
    return out.cpu().numpy(), ts.cpu().numpy()


def surface_intersect(surface, rays, material, device: int = 0):
    """Surface.get_intersect (raytrace.py:1331-1337, 1398-1403, 1479-1516, 1580-1584) as a one-surface device trace."""
    rays = np.atleast_2d(np.asarray(rays, dtype=np.float64))
    rays = np.ascontiguousarray(rays)
    uniq = engine.choose_wavelength_table([material, material], rays)
    packed = engine.pack_system([surface], [material, material], uniq)
    mode, idx, n_out = engine.resolve_keep([1], packed.n_slabs)
    opts = engine.make_opts(mode, idx, "f64")
    opts.flags = _ffi.FLAG_INTERSECT_ONLY
    rays = np.ascontiguousarray(rays)
    out = np.empty((1, rays.shape[0], 8))
    _ffi.require_device()
    rc = _ffi.lib().rtb_trace_host(C.byref(packed.sys), rays.ctypes.data, rays.shape[0], out.ctypes.data,
                                   C.byref(opts), device)
    _ffi.check(rc)
    return out[0]
