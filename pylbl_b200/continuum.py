"""MT-CKD continua on the GPU: host-side mirror of pyLBL's continuum plugins.

The reference attaches to every gas the continua registered under its name -- "H2OForeign" and
"H2OSelf" for H2O, else the formula itself (pyLBL/spectroscopy.py:58-65, pyLBL/plugins.py:9-15:
CO2, N2, O2, O3) -- and adds ``continuum.spectra(T, p, vmr, grid)`` per layer into mechanism 1
(spectroscopy.py:194-198).  ``Continuum`` does the same for all layers in one call on the device
(include/pylbl_b200.h, lbl_continuum_*), on the grid (v0, vn, n_per_v) the lines calls use.
The coefficients are the reference's own table (pyLBL/mt_ckd/mt-ckd.nc), converted once to
pylbl_b200/data/mt_ckd.npz by tools/convert_mt_ckd.py.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from pathlib import Path

import numpy as np

from . import _lib
from .gas_optics import default_device, grid_to_ints

TABLE = Path(__file__).resolve().parent / "data" / "mt_ckd.npz"
NAMES = ("CO2", "H2OForeign", "H2OSelf", "N2", "O2", "O3")
_STATE = ("H2O", "CO2", "O3", "N2", "O2")     # the mole fractions the band formulas read


def continua_of(formula):
    """Continuum names the reference driver attaches to a gas (spectroscopy.py:58-65)."""
    if formula == "H2O":
        return ["H2OForeign", "H2OSelf"]
    return [formula] if formula in NAMES else []


class Continuum(object):
    """The MT-CKD coefficient table on one CUDA device."""

    def __init__(self, device=None, table=TABLE):
        self.device = default_device() if device is None else int(device)
        self.ptr = c_void_p()
        lib = _lib.library()
        z = np.load(table)
        lib.lbl_continuum_create(self.device, ctypes.byref(self.ptr))
        for key in z.files:
            if key.endswith("__grid"):
                continue
            data = np.ascontiguousarray(z[key], dtype=np.float64)
            lower, upper, resolution = (float(x) for x in z[key + "__grid"])
            lib.lbl_continuum_set_spectrum(self.ptr, key.encode(), lower, upper, resolution,
                                           int(data.size), data)
        lib.lbl_continuum_finalize(self.ptr)

    def close(self):
        ptr, self.ptr = self.ptr, None
        if ptr is not None and ptr.value:
            _lib.library().lbl_continuum_close(ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_ms(self):
        """(band formulas, interpolate-and-add) kernel durations of the last call [ms]."""
        a, b = ctypes.c_float(0.), ctypes.c_float(0.)
        _lib.library().lbl_continuum_last_ms(self.ptr, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    @staticmethod
    def state(volume_mixing_ratio, n_layers):
        """(n_layers, 6): H2O, CO2, O3, N2, O2 and the sum over all gases -- what the band
        formulas read of the reference's ``vmr`` dictionary.  A gas the atmosphere does not
        hold raises KeyError only when a formula needs it (as the reference's lookup would);
        here it is entered as NaN so that such a continuum comes out NaN rather than wrong."""
        out = np.full((n_layers, 6), np.nan)
        for k, name in enumerate(_STATE):
            if name in volume_mixing_ratio:
                out[:, k] = np.asarray(volume_mixing_ratio[name], dtype=np.float64).ravel()
        out[:, 5] = sum(np.asarray(x, dtype=np.float64).ravel() for x in volume_mixing_ratio.values())
        return out

    def spectra(self, name, temperature, pressure, volume_mixing_ratio, grid=None, bounds=None,
                out=None, mix=None, row0=0):
        """Continuum extinction [m-1] of continuum ``name`` (or the sum of several: a list of names,
        one pass over the output) for every layer, shape
        (n_layers, (vn-v0)*n_per_v): row L is ``BandedContinuum.spectra(T[L], p[L], vmr[L], grid)``
        (pyLBL/mt_ckd/utils.py:157-174).  ``volume_mixing_ratio``: {formula: array over layers}
        of ALL gases of the atmosphere.  ``mix``: an lbl_mix accumulator to add into instead."""
        v0, vn, n_per_v = bounds if bounds is not None else grid_to_ints(grid)
        t = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
        p = np.ascontiguousarray(pressure, dtype=np.float64).ravel()
        if "H2O" not in volume_mixing_ratio:
            raise KeyError("H2O")      # dry_air_number_density, mt_ckd/utils.py:44
        names = [name] if isinstance(name, str) else list(name)
        for one in names:
            for gas in {"CO2": ["CO2"], "O3": ["O3"], "N2": ["N2", "O2"], "O2": ["O2", "N2"]}.get(one, []):
                if gas not in volume_mixing_ratio:
                    raise KeyError(gas)
        name = ",".join(names)
        state = np.ascontiguousarray(self.state(volume_mixing_ratio, t.size))
        n = (vn - v0) * n_per_v
        lib = _lib.library()
        if mix is None:
            if out is None:
                out = np.empty((t.size, n))
            if out.shape != (t.size, n) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous float64 array of shape (n_layers, n)")
        for lo in range(0, t.size, self.MAX_LAYERS):       # layers are a grid dimension of the kernels
            hi = min(lo + self.MAX_LAYERS, t.size)
            lib.lbl_continuum_compute(self.ptr, name.encode(), hi - lo, t[lo:hi], p[lo:hi], state[lo:hi],
                                      v0, vn, n_per_v, mix, int(row0) + lo if mix is not None else 0,
                                      None if mix is not None else out[lo:hi].ctypes.data_as(c_void_p))
        return out if mix is None else None

    MAX_LAYERS = 32768
