import numpy as np


def relerr(a, b):
    """max relative error with the denominator max(|ref|, 1e-300); NaN positions must coincide."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"NaN positions differ: {int((na != nb).sum())} entries"
    ok = ~na
    if not ok.any():
        return 0.0
    d = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
    return float(d.max())


def abs_rel_err(a, b, scale):
    """max |a-b| / scale (for quantities that pass through zero, e.g. transverse errors)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"NaN positions differ: {int((na != nb).sum())} entries"
    ok = ~na
    if not ok.any():
        return 0.0
    return float(np.abs(a[ok] - b[ok]).max() / scale)


def n_bits_differ(a, b):
    """number of entries whose bit patterns differ; NaN == NaN regardless of sign/payload"""
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    return int(((a.view(np.uint64) != b.view(np.uint64)) & ~both_nan).sum())


def bits_equal(a, b):
    return n_bits_differ(a, b) == 0
