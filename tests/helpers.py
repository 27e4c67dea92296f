"""Parity metrics shared by the tests."""
import numpy as np
from scipy.ndimage import maximum_filter1d

FP64_TOL = 1.0e-9   # BASELINE.json north_star: relative error <= 1e-9 in FP64 mode
FP32_TOL = 1.0e-4   # stated tolerance of the opt-in FP32 mode


def scaled_error(k, k_ref, n_per_v, cut_off=25):
    """max |k - k_ref| / max(|k_ref| within +-cut_off cm-1)  (SURVEY.md section 8(d)).

    With the pedestal removed the reference spectrum crosses zero, so a pointwise relative
    error is ill-posed; the local scale is the largest reference value in the point's own
    line window.
    """
    a = np.abs(k_ref)
    scale = maximum_filter1d(a, size=2 * cut_off * n_per_v + 1, mode="nearest")
    scale = np.maximum(scale, np.finfo(float).tiny)
    return float(np.max(np.abs(k - k_ref) / scale))


def relative_error(k, k_ref):
    """Pointwise max |k - k_ref| / |k_ref| over points where the reference is non-zero."""
    nz = k_ref != 0
    if not np.any(nz):
        return float(np.max(np.abs(k)))
    return float(np.max(np.abs(k[nz] - k_ref[nz]) / np.abs(k_ref[nz])))
