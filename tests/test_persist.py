"""Sweep results on disk in the reference script's zarr layout (scripts/2024_04_01_lightsheet.py:51-61), CPU only."""
import json

import numpy as np
import pytest

from ray_trace_pb_b200 import persist


def _histories(n_cfg=4, n_slabs=7, n_rays=11, seed=0):
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(n_cfg, n_slabs, n_rays, 8))
    h[1, 3:, 2] = np.nan            # a ray that died
    h[-1, :, :, 3] = -0.0
    return h


def test_script_style_round_trip(tmp_path):
    h = _histories()
    z = persist.open(tmp_path / "rays.zarr", "w")
    z.create("rays", shape=h.shape, chunks=(1,) + h.shape[1:], dtype=float)
    z.array("radius_curvatures", np.linspace(5, 55, 4), dtype=float)
    z.rays.attrs["array_columns"] = persist.ARRAY_COLUMNS
    z.attrs["settings"] = {"nrays": 11, "wavelength": 0.561, "n_immersion": np.float64(1.333)}
    for ii in (0, 1, 3):
        z.rays[ii] = h[ii]

    r = persist.open(tmp_path / "rays.zarr")
    assert r.array_keys() == ["radius_curvatures", "rays"]
    assert r.rays.shape == h.shape and r.rays.chunks == (1,) + h.shape[1:] and r.rays.dtype == np.float64
    assert r.rays.attrs["array_columns"] == ["x", "y", "z", "dx", "dy", "dz", "phase", "wavelength"]
    assert r.attrs["settings"]["n_immersion"] == 1.333
    for ii in (0, 1, 3):
        assert np.array_equal(r.rays[ii].view(np.uint64), h[ii].view(np.uint64))      # NaNs and -0.0 included
    assert np.array_equal(r.rays[2], np.zeros(h.shape[1:]))                            # never written: fill value
    assert np.array_equal(r["radius_curvatures"][...], np.linspace(5, 55, 4))
    full = np.asarray(r.rays)
    assert full.shape == h.shape and np.array_equal(full[3], h[3])
    assert np.array_equal(r.rays[-1], h[3])
    with pytest.raises(PermissionError):
        r.rays[0] = h[0]
    with pytest.raises(PermissionError):
        r.attrs["x"] = 1
    with pytest.raises(IndexError):
        r.rays[4]


def test_on_disk_format_is_zarr_v2(tmp_path):
    h = _histories(n_cfg=2)
    persist.save_sweep(tmp_path / "s.zarr", h, parameters={"defocus": [0.0, 1.0]}, settings={"a": 1})
    root = tmp_path / "s.zarr"
    assert json.loads((root / ".zgroup").read_text()) == {"zarr_format": 2}
    meta = json.loads((root / "rays" / ".zarray").read_text())
    assert meta == {"chunks": [1, 7, 11, 8], "compressor": None, "dtype": "<f8", "fill_value": 0.0, "filters": None,
                    "order": "C", "shape": [2, 7, 11, 8], "zarr_format": 2}
    # one raw little-endian C-order file per configuration, named by its chunk index
    raw = np.fromfile(root / "rays" / "1.0.0.0", dtype="<f8").reshape(7, 11, 8)
    assert np.array_equal(raw.view(np.uint64), h[1].view(np.uint64))
    assert sorted(p.name for p in (root / "rays").iterdir()) == [".zarray", ".zattrs", "0.0.0.0", "1.0.0.0"]
    assert json.loads((root / "rays" / ".zattrs").read_text())["array_columns"] == persist.ARRAY_COLUMNS
    assert json.loads((root / ".zattrs").read_text()) == {"settings": {"a": 1}}
    assert json.loads((root / "defocus" / ".zarray").read_text())["shape"] == [2]


def test_multi_row_chunks_and_whole_array_assignment(tmp_path):
    z = persist.open(tmp_path / "g.zarr", "w")
    a = z.create("a", shape=(10, 3), chunks=(4, 3), dtype=np.float32, fill_value=float("nan"))
    data = np.arange(30, dtype=np.float32).reshape(10, 3)
    a[...] = data
    assert sorted(p.name for p in (tmp_path / "g.zarr" / "a").iterdir() if not p.name.startswith(".")) == \
        ["0.0", "1.0", "2.0"]
    assert (tmp_path / "g.zarr" / "a" / "2.0").stat().st_size == 4 * 3 * 4          # edge chunk stored at full size
    assert np.array_equal(np.asarray(z.a), data)
    a[5] = [-1, -2, -3]
    assert np.array_equal(z.a[5], [-1, -2, -3]) and np.array_equal(z.a[4], data[4])
    assert json.loads((tmp_path / "g.zarr" / "a" / ".zarray").read_text())["fill_value"] == "NaN"
    s = z.array("scalar", 2.5)
    assert s.shape == () and float(z.scalar[...]) == 2.5
    # re-creating an array drops its old chunks
    z.create("a", shape=(2, 3), chunks=(1, 3))
    assert np.array_equal(np.asarray(z.a), np.zeros((2, 3)))
    with pytest.raises(NotImplementedError):
        z.create("b", shape=(4, 4), chunks=(2, 2))[0] = np.zeros(4)


def test_errors_and_append_mode(tmp_path):
    with pytest.raises(FileNotFoundError):
        persist.open(tmp_path / "missing.zarr")
    with pytest.raises(ValueError):
        persist.open(tmp_path / "x.zarr", "x")
    with pytest.raises(ValueError):
        persist.save_sweep(tmp_path / "e.zarr", [np.zeros((3, 8))])
    with pytest.raises(ValueError):
        persist.save_sweep(tmp_path / "e.zarr", [np.zeros((3, 4, 8)), np.zeros((3, 5, 8))])
    z = persist.open(tmp_path / "ap.zarr", "a")
    z.array("p", [1.0, 2.0])
    z2 = persist.open(tmp_path / "ap.zarr", "a")
    z2.p[1] = 5.0
    assert np.array_equal(persist.open(tmp_path / "ap.zarr").p[...], [1.0, 5.0])
    assert "p" in z2 and "q" not in z2
    with pytest.raises(AttributeError):
        z2.q
    (tmp_path / "ap.zarr" / "p" / ".zarray").write_text(json.dumps(
        {"chunks": [2], "compressor": {"id": "blosc"}, "dtype": "<f8", "fill_value": 0, "filters": None, "order": "C",
         "shape": [2], "zarr_format": 2}))
    with pytest.raises(NotImplementedError):
        z2.p


def test_reads_what_the_zarr_package_reads(tmp_path):
    zarr = pytest.importorskip("zarr")
    h = _histories(n_cfg=2)
    persist.save_sweep(tmp_path / "s.zarr", h, parameters={"defocus": [0.0, 1.0]})
    z = zarr.open(str(tmp_path / "s.zarr"), "r")
    assert np.array_equal(np.asarray(z["rays"][1]).view(np.uint64), h[1].view(np.uint64))
    assert list(z["rays"].attrs["array_columns"]) == persist.ARRAY_COLUMNS
