"""
ctypes binding of librtb.so (C ABI in include/rtb.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ray_trace_pb_b200/csrc`` into
``ray_trace_pb_b200/_lib/librtb.so``.  There is no CPU fallback: if the library is missing, or no CUDA device is
visible when a trace is requested, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import weakref
from pathlib import Path

import numpy as np

import os

# RTB_LIBRARY_PATH selects another build of the same library (used to A/B kernel variants on the GPU box)
LIB_PATH = Path(os.environ.get("RTB_LIBRARY_PATH") or (Path(__file__).resolve().parent / "_lib" / "librtb.so"))

RTB_ABI_VERSION = 5
RTB_COMM_ID_BYTES = 128
RTB_MAX_SURFACES = 64
RTB_MAX_WAVELENGTHS = 8
RTB_N_STATS = 12

RTB_OK = 0
RTB_ERR_INVALID = -1
RTB_ERR_UNSUPPORTED = -2
RTB_ERR_CUDA = -3
RTB_ERR_NOMEM = -4

SURF_FLAT, SURF_SPHERE, SURF_MIRROR, SURF_PERFECT_LENS = 0, 1, 2, 3
MAT_CONSTANT, MAT_SELLMEIER, MAT_TABLE_ONLY = 0, 1, 2
F64_EXACT, F32_FAST, F64_FAST = 0, 1, 2
KEEP_ALL, KEEP_LAST, KEEP_LIST, KEEP_NONE = 0, 1, 2, 3
SRC_COLLIMATED, SRC_FAN, SRC_GRID = 0, 1, 2
HINT_DEGENERATE = 1          # rtb_surface.hints
FLAG_INTERSECT_ONLY = 1
FLAG_PLANES_IN = 2
FLAG_PLANES_OUT = 4

_d3 = C.c_double * 3


class RtbSurface(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hints", C.c_int32),
                ("center", _d3), ("normal", _d3), ("input_axis", _d3),
                ("radius", C.c_double), ("radius_sq", C.c_double), ("abs_radius", C.c_double),
                ("aperture_rad", C.c_double), ("focal_len", C.c_double), ("normal_f", _d3),
                ("sin_alpha", C.c_double)]


class RtbMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("b", _d3), ("c", _d3), ("n_const", C.c_double)]


class RtbSystem(C.Structure):
    _fields_ = [("n_surfaces", C.c_int32), ("n_wavelengths", C.c_int32),
                ("surfaces", C.POINTER(RtbSurface)), ("materials", C.POINTER(RtbMaterial)),
                ("wavelengths", C.POINTER(C.c_double)), ("n_table", C.POINTER(C.c_double))]


class RtbReduce(C.Structure):
    _fields_ = [("slab", C.c_int32), ("grid_n", C.c_int32), ("origin", _d3), ("e1", _d3), ("e2", _d3),
                ("phase_ref", C.c_double), ("grid_half_width", C.c_double),
                ("stats_dev", C.c_void_p), ("grid_dev", C.c_void_p)]


class RtbTraceOpts(C.Structure):
    _fields_ = [("precision", C.c_int32), ("keep_mode", C.c_int32), ("n_keep", C.c_int32), ("flags", C.c_int32),
                ("keep_slabs", C.POINTER(C.c_int32)), ("reduce", C.POINTER(RtbReduce))]


class RtbSource(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("n_a", C.c_int64), ("n_b", C.c_int64),
                ("a_max", C.c_double), ("b_max", C.c_double), ("b_start", C.c_double),
                ("pt", _d3), ("axis", _d3), ("e1", _d3), ("e2", _d3), ("wavelength", C.c_double)]


class RtbError(RuntimeError):
    pass


_lib = None

# every symbol include/rtb.h declares (tests check the .so exports all of them)
EXPORTS = ["rtb_abi_version", "rtb_last_error", "rtb_device_count", "rtb_launch_count", "rtb_pure_launch_count", "rtb_trace_device",
           "rtb_trace_host", "rtb_trace_source", "rtb_trace_sources", "rtb_generate_device", "rtb_reduce_init",
           "rtb_intersect_rays_device", "rtb_measure_dfma_rate", "rtb_measure_dfma_chain_rate",
           "rtb_measure_copy_bandwidth", "rtb_ray2plane_device", "rtb_distinct_wavelengths_device", "rtb_distinct_wavelengths_host", "rtb_host_alloc", "rtb_host_free",
           "rtb_selftest_exact_math", "rtb_psf_scratch_doubles", "rtb_psf_from_grid_device", "rtb_tune", "rtb_last_probe_counts",
           "rtb_comm_available", "rtb_comm_unique_id", "rtb_comm_init", "rtb_comm_size", "rtb_comm_allreduce_grid",
           "rtb_comm_allreduce_stats", "rtb_comm_destroy"]


def lib():
    """Load librtb.so once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RtbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C ray_trace_pb_b200/csrc`. There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i64, i32, dp = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double)
    L.rtb_abi_version.restype = i32
    L.rtb_last_error.restype = C.c_char_p
    L.rtb_device_count.restype = i32
    L.rtb_launch_count.restype = i64
    L.rtb_pure_launch_count.restype = i64
    L.rtb_trace_device.argtypes = [C.POINTER(RtbSystem), vp, i64, vp, C.POINTER(RtbTraceOpts), i32, vp]
    L.rtb_trace_host.argtypes = [C.POINTER(RtbSystem), vp, i64, vp, C.POINTER(RtbTraceOpts), i32]
    L.rtb_trace_source.argtypes = [C.POINTER(RtbSystem), C.POINTER(RtbSource), i64, i64, vp,
                                   C.POINTER(RtbTraceOpts), i32, vp]
    L.rtb_trace_sources.argtypes = [C.POINTER(RtbSystem), C.POINTER(RtbSource), i32, i64, i64, vp,
                                    C.POINTER(RtbTraceOpts), i32, vp]
    L.rtb_generate_device.argtypes = [C.POINTER(RtbSource), i64, i64, vp, i32, vp]
    L.rtb_reduce_init.argtypes = [C.POINTER(RtbReduce), i32, vp]
    L.rtb_intersect_rays_device.argtypes = [vp, i64, vp, i64, vp, i32, vp]
    L.rtb_ray2plane_device.argtypes = [vp, i64, vp, i64, vp, i64, vp, i32, vp, vp, i32, vp]
    L.rtb_distinct_wavelengths_device.argtypes = [vp, i64, vp, dp, C.POINTER(C.c_int32), i32, vp]
    L.rtb_distinct_wavelengths_host.argtypes = [vp, i64, dp, C.POINTER(C.c_int32)]
    L.rtb_distinct_wavelengths_host.restype = i32
    L.rtb_selftest_exact_math.argtypes = [i32, C.c_uint64, i64, C.POINTER(C.c_uint64)]
    L.rtb_selftest_exact_math.restype = i32
    L.rtb_psf_scratch_doubles.argtypes = [i32, i32, i32]
    L.rtb_psf_scratch_doubles.restype = i64
    L.rtb_psf_from_grid_device.argtypes = [vp, i32, C.c_double, i32, C.c_double, i32, vp, i64, vp, vp, vp, i32, vp]
    L.rtb_psf_from_grid_device.restype = i32
    L.rtb_measure_dfma_rate.argtypes = [i32, dp, dp]
    L.rtb_measure_dfma_chain_rate.argtypes = [i32, i32, dp, dp]
    L.rtb_measure_copy_bandwidth.argtypes = [i32, i64, dp]
    L.rtb_last_probe_counts.argtypes = [C.POINTER(C.c_uint32), i32]
    L.rtb_last_probe_counts.restype = i32
    L.rtb_comm_available.restype = i32
    L.rtb_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.rtb_comm_unique_id.restype = i32
    L.rtb_comm_init.argtypes = [C.POINTER(vp), i32, i32, vp, i32]
    L.rtb_comm_init.restype = i32
    L.rtb_comm_size.argtypes = [vp]
    L.rtb_comm_size.restype = i32
    L.rtb_comm_allreduce_grid.argtypes = [vp, vp, i64, vp]
    L.rtb_comm_allreduce_grid.restype = i32
    L.rtb_comm_allreduce_stats.argtypes = [vp, vp, i32, vp]
    L.rtb_comm_allreduce_stats.restype = i32
    L.rtb_comm_destroy.argtypes = [vp]
    L.rtb_comm_destroy.restype = i32
    L.rtb_tune.argtypes = [C.c_char_p, i64]
    L.rtb_tune.restype = i32
    L.rtb_host_alloc.argtypes = [C.c_size_t]
    L.rtb_host_alloc.restype = vp
    L.rtb_host_free.argtypes = [vp]
    L.rtb_host_free.restype = None
    for name in ("rtb_trace_device", "rtb_trace_host", "rtb_trace_source", "rtb_trace_sources", "rtb_generate_device",
                 "rtb_reduce_init",
                 "rtb_intersect_rays_device", "rtb_measure_dfma_rate", "rtb_measure_dfma_chain_rate",
                 "rtb_measure_copy_bandwidth", "rtb_ray2plane_device", "rtb_distinct_wavelengths_device"):
        getattr(L, name).restype = i32
    if L.rtb_abi_version() != RTB_ABI_VERSION:
        raise RtbError(f"librtb.so ABI {L.rtb_abi_version()} != binding ABI {RTB_ABI_VERSION}; rebuild the library")
    _lib = L
    return L


def check(rc: int):
    """Map a negative rtb_status to the Python exception the reference API would raise for that class of error."""
    if rc == RTB_OK:
        return
    msg = lib().rtb_last_error().decode("utf-8", "replace")
    if rc == RTB_ERR_INVALID:
        raise ValueError(msg)
    if rc == RTB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == RTB_ERR_NOMEM:
        raise MemoryError(msg)
    raise RtbError(msg)


def require_device() -> int:
    n = lib().rtb_device_count()
    if n <= 0:
        raise RtbError("no CUDA device visible: ray_trace_pb_b200 traces only on the GPU (no CPU fallback)")
    return n


class _PinnedPool:
    """
    Page-locked host buffers recycled across calls.  Pinning memory costs ~0.45 s/GiB and a fresh pageable result
    array costs a page fault per 4 KiB, which is what bounds a drop-in ``System.ray_trace`` returning a GiB-sized
    history; a recycled pinned buffer is written by DMA at PCIe speed instead.  Buffers return to the pool when the
    NumPy arrays viewing them are garbage collected.  RTB_PINNED_POOL_GB caps the idle bytes kept (default 8).
    """

    GRANULE = 32 << 20

    def __init__(self):
        self.free = {}          # rounded size -> [ptr, ...]
        self.idle_bytes = 0
        self.cap = int(float(os.environ.get("RTB_PINNED_POOL_GB", "8")) * (1 << 30))

    def take(self, nbytes: int):
        size = max(self.GRANULE, -(-nbytes // self.GRANULE) * self.GRANULE)
        bucket = self.free.get(size)
        if bucket:
            self.idle_bytes -= size
            return bucket.pop(), size
        ptr = lib().rtb_host_alloc(size)
        if not ptr:
            self.trim(0)
            ptr = lib().rtb_host_alloc(size)
            if not ptr:
                check(RTB_ERR_NOMEM)
        return ptr, size

    def give(self, ptr, size):
        if self.idle_bytes + size > self.cap:
            _free_pinned(ptr)
            return
        self.free.setdefault(size, []).append(ptr)
        self.idle_bytes += size

    def trim(self, keep_bytes: int = 0):
        for size, bucket in list(self.free.items()):
            while bucket and self.idle_bytes > keep_bytes:
                _free_pinned(bucket.pop())
                self.idle_bytes -= size


_pool = _PinnedPool()


def pinned_empty(shape, dtype=np.float64, pooled: bool = False) -> np.ndarray:
    """A NumPy array in page-locked memory (released -- or returned to the pool -- when it is garbage collected)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = count * dtype.itemsize
    if pooled:
        ptr, size = _pool.take(max(nbytes, 1))
        buf = (C.c_char * size).from_address(ptr)
        weakref.finalize(buf, _pool.give, ptr, size)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    ptr = lib().rtb_host_alloc(max(nbytes, 1))
    if not ptr:
        check(RTB_ERR_NOMEM)
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
    # every view of the returned array reaches `buf` through its .base chain, so the allocation lives as long as
    # any of them does
    weakref.finalize(buf, _free_pinned, ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def result_array(shape) -> np.ndarray:
    """Where a host-path trace puts its result: recycled pinned memory for big histories, plain NumPy for small."""
    nbytes = int(np.prod(shape)) * 8
    if nbytes >= (8 << 20):
        try:
            return pinned_empty(shape, pooled=True)
        except MemoryError:
            pass        # no page-locked memory left: a pageable result works too (rtb_trace_host stages it), only slower
    return np.empty(shape, dtype=np.float64)


def _free_pinned(ptr):
    try:
        lib().rtb_host_free(ptr)
    except Exception:
        pass
