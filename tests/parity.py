"""Comparison helpers for the parity tests: bit-level equality with all NaNs treated as one value."""
from __future__ import annotations

import hashlib

import numpy as np


def canonical(a: np.ndarray) -> np.ndarray:
    """uint64 view with every NaN replaced by one canonical quiet-NaN pattern (sign of zero is kept)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    bits = a.view(np.uint64).copy()
    bits[np.isnan(a)] = np.uint64(0x7FF8000000000000)
    return bits


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(canonical(a).tobytes()).hexdigest()


def mismatch_report(got: np.ndarray, want: np.ndarray, limit: int = 8) -> str:
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if got.shape != want.shape:
        return f"shape {got.shape} != {want.shape}"
    bad = np.argwhere(canonical(got) != canonical(want))
    lines = [f"{len(bad)} of {got.size} entries differ"]
    nan_flip = np.sum(np.isnan(got) != np.isnan(want))
    lines.append(f"NaN-mask flips: {int(nan_flip)}")
    for idx in bad[:limit]:
        idx = tuple(idx)
        lines.append(f"  at {idx}: got {got[idx]!r} ({got[idx].hex() if np.isfinite(got[idx]) else got[idx]}) "
                     f"want {want[idx]!r} ({want[idx].hex() if np.isfinite(want[idx]) else want[idx]})")
    return "\n".join(lines)


def assert_bit_identical(got, want, what: str = ""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if not np.array_equal(canonical(got), canonical(want)):
        raise AssertionError(f"{what}: not bit-identical\n" + mismatch_report(got, want))


def assert_close_same_mask(got, want, rtol: float, what: str = "", atol: float = 0.0):
    """Identical NaN masks; finite entries within rtol relative (plus atol absolute)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    flips = np.isnan(got) != np.isnan(want)
    assert not flips.any(), f"{what}: {int(flips.sum())} NaN-mask flips\n" + mismatch_report(got, want)
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok])
    tol = atol + rtol * np.abs(want[ok])
    worst = np.max(err - tol) if err.size else 0.0
    assert worst <= 0, f"{what}: max excess over tolerance {worst:.3e} (rtol {rtol:g})"
