"""
Ray histories on disk in the layout the reference's sweep scripts use (scripts/2024_04_01_lightsheet.py:51-61):
a zarr group with an array ``rays`` of shape ``(n_cfg, n_slabs, n_rays, 8)``, one chunk per configuration,
``attrs["array_columns"]`` naming the 8 columns, plus small 1-D parameter arrays and a ``settings`` attribute.

The ``zarr`` package is not a dependency of this package (and is absent from the build image), so this module writes
and reads the zarr **v2 directory-store format** directly -- ``.zgroup`` / ``.zarray`` / ``.zattrs`` JSON and one raw
little-endian C-order file per chunk (``compressor: null``), which ``zarr.open(path)`` reads as is.  Only what the
sweep scripts need is implemented: arrays are written and read in whole chunks along the leading axis.

    z = persist.open(fdir / "rays.zarr", "w")
    z.create("rays", shape=(n_cfg, n_slabs, n_rays, 8), chunks=(1, n_slabs, n_rays, 8), dtype=float)
    z.array("radius_curvatures", rad_curvs)
    z.rays.attrs["array_columns"] = persist.ARRAY_COLUMNS
    z.attrs["settings"] = settings
    for ii, system in enumerate(systems):
        z.rays[ii] = system.ray_trace(rays, m_in, m_out)          # one chunk file per configuration
"""
from __future__ import annotations

import json
import math
import os
from pathlib import Path

import numpy as np

ARRAY_COLUMNS = ["x", "y", "z", "dx", "dy", "dz", "phase", "wavelength"]


def _json_default(o):
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, Path):
        return str(o)
    raise TypeError(f"{type(o).__name__} is not JSON serialisable")


class Attrs:
    """dict-like view of a node's ``.zattrs`` file; every assignment is written through"""

    def __init__(self, path: Path, writable: bool):
        self._path = path
        self._writable = writable

    def _load(self) -> dict:
        if self._path.exists():
            return json.loads(self._path.read_text())
        return {}

    def __getitem__(self, key):
        return self._load()[key]

    def __contains__(self, key):
        return key in self._load()

    def __setitem__(self, key, value):
        if not self._writable:
            raise PermissionError("store opened read-only")
        d = self._load()
        d[key] = value
        self._path.write_text(json.dumps(d, indent=4, default=_json_default))

    def asdict(self) -> dict:
        return self._load()

    def keys(self):
        return self._load().keys()


class Array:
    """one zarr v2 array stored as uncompressed chunks; whole-chunk access along the leading axis"""

    def __init__(self, path: Path, writable: bool):
        self._path = path
        self._writable = writable
        meta = json.loads((path / ".zarray").read_text())
        if meta.get("zarr_format") != 2:
            raise ValueError(f"{path}: zarr_format {meta.get('zarr_format')} is not supported")
        if meta.get("compressor") is not None or meta.get("filters"):
            raise NotImplementedError(f"{path}: compressed / filtered chunks need the zarr package")
        if meta.get("order", "C") != "C":
            raise NotImplementedError(f"{path}: only C-order chunks are supported")
        self.shape = tuple(meta["shape"])
        self.chunks = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        fill = meta.get("fill_value", 0)
        self.fill_value = float(fill) if isinstance(fill, str) else (0 if fill is None else fill)
        self._sep = meta.get("dimension_separator", ".")
        self.attrs = Attrs(path / ".zattrs", writable)

    # -- chunk grid -------------------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    def _grid(self):
        return tuple(math.ceil(s / c) if c else 0 for s, c in zip(self.shape, self.chunks))

    def _chunk_file(self, idx) -> Path:
        return self._path / (self._sep.join(str(i) for i in idx) if idx else "0")

    def _read_chunk(self, idx) -> np.ndarray:
        f = self._chunk_file(idx)
        if not f.exists():
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        return np.fromfile(f, dtype=self.dtype).reshape(self.chunks)

    def _write_chunk(self, idx, block: np.ndarray):
        tmp = self._chunk_file(idx).with_suffix(".partial")
        np.ascontiguousarray(block, dtype=self.dtype).tofile(tmp)
        os.replace(tmp, self._chunk_file(idx))

    def _leading_only(self):
        if any(c != s for c, s in zip(self.chunks[1:], self.shape[1:])):
            raise NotImplementedError("only arrays chunked along their leading axis are supported")

    # -- access ----------------------------------------------------------------------------------------------
    def __len__(self):
        return self.shape[0]

    def __setitem__(self, key, value):
        if not self._writable:
            raise PermissionError("store opened read-only")
        self._leading_only()
        value = np.asarray(value)
        if isinstance(key, (int, np.integer)):
            lo, hi = int(key), int(key) + 1
            if lo < 0:
                lo, hi = lo + self.shape[0], hi + self.shape[0]
            value = value.reshape((1,) + self.shape[1:])
        elif key is Ellipsis or (isinstance(key, slice) and key == slice(None)):
            lo, hi = 0, self.shape[0]
            value = np.broadcast_to(value, self.shape)
        else:
            raise NotImplementedError("assign one leading index or the whole array")
        if not 0 <= lo < hi <= self.shape[0]:
            raise IndexError(f"index {key} out of range for axis 0 of size {self.shape[0]}")
        c0 = self.chunks[0]
        for c in range(lo // c0, (hi - 1) // c0 + 1):
            a, b = c * c0, min((c + 1) * c0, self.shape[0])
            idx = (c,) + (0,) * (self.ndim - 1)
            if lo <= a and b <= hi and b - a == c0:
                block = value[a - lo:b - lo]
            else:           # partial chunk: read, patch, write (edge chunks are stored at full chunk size)
                block = self._read_chunk(idx).copy()
                s, e = max(a, lo), min(b, hi)
                block[s - a:e - a] = value[s - lo:e - lo]
            self._write_chunk(idx, block)

    def __getitem__(self, key):
        self._leading_only()
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            if not 0 <= k < self.shape[0]:
                raise IndexError(f"index {key} out of range for axis 0 of size {self.shape[0]}")
            c0 = self.chunks[0]
            return self._read_chunk((k // c0,) + (0,) * (self.ndim - 1))[k % c0].copy()
        if key is Ellipsis or (isinstance(key, slice) and key == slice(None)):
            if self.ndim == 0:
                return self._read_chunk(())
            out = np.empty(self.shape, dtype=self.dtype)
            c0 = self.chunks[0]
            for c in range(self._grid()[0]):
                a, b = c * c0, min((c + 1) * c0, self.shape[0])
                out[a:b] = self._read_chunk((c,) + (0,) * (self.ndim - 1))[:b - a]
            return out
        return self[...][key]

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a.astype(dtype) if dtype is not None else a


class Group:
    """a zarr v2 group: named arrays, attributes"""

    def __init__(self, path: Path, writable: bool):
        self._path = path
        self._writable = writable
        self.attrs = Attrs(path / ".zattrs", writable)

    def create(self, name: str, shape, chunks=None, dtype=float, fill_value=0.0) -> Array:
        if not self._writable:
            raise PermissionError("store opened read-only")
        shape = tuple(int(s) for s in (shape if np.iterable(shape) else (shape,)))
        if chunks is None or chunks is True:
            chunks = shape
        chunks = tuple(int(c) for c in (chunks if np.iterable(chunks) else (chunks,)))
        if len(chunks) != len(shape):
            raise ValueError(f"chunks {chunks} do not match shape {shape}")
        dt = np.dtype(dtype)
        if dt.byteorder == ">":
            raise NotImplementedError("big-endian dtypes are not supported")
        node = self._path / name
        node.mkdir(parents=True, exist_ok=True)
        for old in node.iterdir():          # a fresh array: no stale chunks
            if old.is_file():
                old.unlink()
        fv = fill_value
        if isinstance(fv, float) and not math.isfinite(fv):
            fv = "NaN" if math.isnan(fv) else ("Infinity" if fv > 0 else "-Infinity")
        meta = {"chunks": list(chunks), "compressor": None, "dtype": dt.str,
                "fill_value": fv, "filters": None, "order": "C", "shape": list(shape), "zarr_format": 2}
        (node / ".zarray").write_text(json.dumps(meta, indent=4))
        return Array(node, True)

    def array(self, name: str, data, dtype=None, chunks=None) -> Array:
        data = np.asarray(data, dtype=dtype)
        arr = self.create(name, data.shape, chunks=chunks, dtype=data.dtype)
        if data.ndim == 0:
            arr._write_chunk((), data.reshape(()))
        elif data.size:
            arr[...] = data
        return arr

    def __getitem__(self, name: str) -> Array:
        node = self._path / name
        if not (node / ".zarray").exists():
            raise KeyError(name)
        return Array(node, self._writable)

    def __getattr__(self, name: str) -> Array:
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __contains__(self, name: str) -> bool:
        return (self._path / name / ".zarray").exists()

    def array_keys(self):
        return sorted(p.name for p in self._path.iterdir() if (p / ".zarray").exists())


def open(path, mode: str = "r") -> Group:
    """``mode`` "w" creates (or empties) the group directory, "a" opens for update, "r" read-only"""
    path = Path(path)
    if mode not in ("r", "a", "w"):
        raise ValueError(f"mode {mode!r}: expected 'r', 'a' or 'w'")
    if mode == "w":
        if path.exists():
            import shutil
            shutil.rmtree(path)
        path.mkdir(parents=True)
        (path / ".zgroup").write_text(json.dumps({"zarr_format": 2}, indent=4))
    elif not (path / ".zgroup").exists():
        if mode == "a":
            path.mkdir(parents=True, exist_ok=True)
            (path / ".zgroup").write_text(json.dumps({"zarr_format": 2}, indent=4))
        else:
            raise FileNotFoundError(f"{path} is not a zarr v2 group")
    return Group(path, mode != "r")


def save_sweep(path, histories, parameters: dict | None = None, settings: dict | None = None) -> Group:
    """
    Write a sweep the way the reference's script does: ``histories`` is an iterable of ``(n_slabs, n_rays, 8)``
    arrays (one per configuration, e.g. straight from ``System.ray_trace``), ``parameters`` maps names to the swept
    1-D arrays, ``settings`` goes to the group attributes.
    """
    histories = list(histories)
    if not histories:
        raise ValueError("no histories to save")
    first = np.asarray(histories[0])
    if first.ndim != 3 or first.shape[-1] != 8:
        raise ValueError(f"history of shape {first.shape}: expected (n_slabs, n_rays, 8)")
    z = open(path, "w")
    rays = z.create("rays", shape=(len(histories),) + first.shape, chunks=(1,) + first.shape, dtype=float)
    for name, values in (parameters or {}).items():
        z.array(name, np.asarray(values, dtype=float))
    rays.attrs["array_columns"] = ARRAY_COLUMNS
    if settings is not None:
        z.attrs["settings"] = settings
    for ii, h in enumerate(histories):
        h = np.asarray(h)
        if h.shape != first.shape:
            raise ValueError(f"history {ii} has shape {h.shape}, the first has {first.shape}")
        rays[ii] = h
    return z
