"""Column-batched mirror of ``pyLBL.Spectroscopy`` for the lines mechanism.

The reference driver (pyLBL/spectroscopy.py:144-206) loops gas-outer, layer-inner and makes one
blocking backend call per (gas, layer), converting each cross-section to an absorption
coefficient on the host, ``beta = n*k[:grid.size]`` with ``n = p*x/(kB*T)`` (:18-29, :181-191).
This class produces the same numbers with one GPU batch per gas (all layers of the atmosphere
at once, every gas in flight together) and, for ``output_format="total"``, with the gas sum
done on the device (``Mixture``).  It needs neither xarray nor SQLAlchemy: the atmosphere is
anything with ``temperature``, ``pressure`` and ``gases`` (a pyLBL ``Atmosphere``, whose
members carry ``.data``, or plain arrays / a dict), and results are numpy arrays in the
reference's layout; ``to_dataset`` wraps them into the reference's xarray ``Dataset`` when
xarray is installed.

Mechanism 0 ("lines") and mechanism 1 ("continuum", the MT-CKD continua on the device,
``continua_backend="mt_ckd"`` as in the reference; ``None`` switches them off) are computed
here; cross-sections (arts-crossfit) are another plugin of the reference and stay zero, as they
do in the reference when that engine has no data for a gas (pyLBL/spectroscopy.py:66-70).
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p

import numpy as np

from . import _lib
from .continuum import Continuum, continua_of
from .gas_optics import Gas, default_device, grid_to_ints
from .mixture import Mixture, number_density

MECHANISMS = ["lines", "continuum", "cross_section"]   # pyLBL/spectroscopy.py:126


def _values(x):
    return np.asarray(getattr(x, "data", x), dtype=np.float64)


def _member(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


class Spectroscopy(object):
    """Line-by-line absorption of a whole atmosphere on the GPU.

    Attributes:
        temperature, pressure: float64 arrays of the atmosphere's shape [K], [Pa].
        gases: {formula: mole-fraction array of the same shape}.
        grid: wavenumber grid [cm-1] (integer first point, step 1/integer: README.rst:229-230).
    """

    def __init__(self, atmosphere, grid, database, device=None, precision="fp64", cache_dir=None,
                 continua_backend="mt_ckd"):
        self.temperature = _values(_member(atmosphere, "temperature"))
        self.pressure = _values(_member(atmosphere, "pressure"))
        self.gases = {name: _values(x) for name, x in dict(_member(atmosphere, "gases")).items()}
        for name, x in self.gases.items():
            if x.shape != self.temperature.shape:
                raise ValueError(f"mole fraction of {name} does not have the atmosphere's shape")
        self.grid = np.asarray(grid, dtype=np.float64)
        self.database = getattr(database, "path", database)
        self.device = default_device() if device is None else int(device)
        self.precision = precision
        self.cache_dir = cache_dir
        self.cache = {}          # formula -> Gas, or None when the database has no such molecule
        self._pinned = {}        # formula -> PinnedArray (staging of the per-gas formats)
        if continua_backend not in ("mt_ckd", None):
            raise ValueError("continua_backend must be 'mt_ckd' or None")
        self.continua_backend = continua_backend     # pyLBL/spectroscopy.py:89,119
        self._continuum = None

    def _continua(self):
        if self.continua_backend is None:
            return None
        if self._continuum is None:
            self._continuum = Continuum(self.device)
        return self._continuum

    PINNED_LIMIT = 2 << 30

    def _gas(self, name):
        """The reference tolerates molecules the database has no line data for (``gas = None``
        on AliasNotFoundError / IsotopologuesNotFoundError / TipsDataNotFoundError /
        TransitionsNotFoundError, pyLBL/spectroscopy.py:53-57) -- and nothing else: a missing
        database file, a CUDA failure or a corrupt cache still raise.  (A molecule without TIPS
        rows or without transitions opens and yields zeros, pyLBL/c_lib/absorption.c:53-59.)"""
        if name not in self.cache:
            try:
                gas = Gas(self.database, name, devices=[self.device], precision=self.precision,
                          cache_dir=self.cache_dir)
                gas._handle(self.device)
            except ValueError:
                if "not found in database" not in _lib.last_error():
                    raise
                gas = None
            self.cache[name] = gas
        return self.cache[name]

    def _staging(self, name, shape):
        """Page-locked destination of one gas's spectra, kept across calls (device-to-host
        copies into pageable memory are synchronous: the gases would run one after the other).
        Beyond PINNED_LIMIT bytes per gas an ordinary array is used."""
        count = int(np.prod(shape))
        if count * 8 > self.PINNED_LIMIT:
            return np.empty(shape)
        held = self._pinned.get(name)
        if held is None or held.array.size < count:
            self._pinned.pop(name, None)
            held = _lib.PinnedArray((count,))
            self._pinned[name] = held
        return held.array[:count].reshape(shape)

    def close(self):
        for gas in self.cache.values():
            if gas is not None:
                gas.close()
        self.cache = {}
        self._pinned = {}
        if self._continuum is not None:
            self._continuum.close()
            self._continuum = None

    def compute_absorption(self, output_format="all", remove_pedestal=None, cut_off=25):
        """Absorption coefficients [m-1], pyLBL/spectroscopy.py:144-206.

        Args:
            output_format: "all"   {"<gas>_absorption": (*shape, 3, grid.size)}, mechanism axis
                                   as in the reference (only "lines" is filled);
                           "gas"   {"<gas>_absorption": (*shape, grid.size)};
                           "total" {"absorption": (*shape, grid.size)}, summed on the device.
            remove_pedestal: None = whether the continuum backend is MT-CKD, as in the reference
                             (pyLBL/spectroscopy.py:163-164).
        Returns:
            dict of numpy arrays, plus "wavenumber" (and "mechanism" for "all").
        """
        if output_format not in ("all", "gas", "total"):
            raise ValueError("output_format must be 'all', 'gas' or 'total'")
        if remove_pedestal is None:
            remove_pedestal = self.continua_backend == "mt_ckd"
        shape = self.temperature.shape
        t = np.ascontiguousarray(self.temperature.ravel())
        p = np.ascontiguousarray(self.pressure.ravel())
        v0, vn, n_per_v = grid_to_ints(self.grid)
        n_k = (vn - v0) * n_per_v
        size = self.grid.size
        out = {"wavenumber": self.grid}
        present = {name: self._gas(name) for name in self.gases}

        continuum = self._continua()
        flat = {name: np.ascontiguousarray(x.ravel()) for name, x in self.gases.items()}
        if output_format == "total":
            total = np.zeros((t.size, size))
            names = [name for name, gas in present.items() if gas is not None]
            if names or continuum is not None:
                mix = Mixture.from_gases({n: present[n] for n in names}, self.device)
                k = mix.total_absorption(t, p, flat if continuum is not None else {n: flat[n] for n in names},
                                         bounds=(v0, vn, n_per_v), remove_pedestal=remove_pedestal,
                                         cut_off=cut_off, continuum=continuum)
                total = k[:, :size]
                mix.close()
            out["absorption"] = total.reshape(shape + (size,))
            return out

        # Every gas in flight at once: submit all, then collect (the copies of one gas overlap
        # the kernels of the next).
        lib = _lib.library()
        ped = 1 if remove_pedestal else 0
        pending = []
        for name, gas in present.items():
            if gas is None:
                continue
            x = np.ascontiguousarray(self.gases[name].ravel())
            k = self._staging(name, (t.size, n_k))
            h = gas._handle(self.device)
            lib.lbl_gas_submit(h.ptr, t.size, p, t, x, v0, vn, n_per_v, int(cut_off), ped,
                               gas.precision, k.ctypes.data_as(c_void_p))
            pending.append((name, h, x, k))
        lines = {}
        for name, h, x, k in pending:
            lib.lbl_gas_wait(h.ptr)
            lines[name] = number_density(t, p, x)[:, None] * k[:, :size]   # n*k[:grid.size]
        for name in self.gases:
            beta = lines.get(name)
            cont = None
            if continuum is not None:
                for cname in continua_of(name):                   # spectroscopy.py:194-198
                    k = continuum.spectra(cname, t, p, flat, bounds=(v0, vn, n_per_v))[:, :size]
                    cont = k if cont is None else cont + k
            if output_format == "gas":
                arr = np.zeros((t.size, size)) if beta is None else beta
                if cont is not None:
                    arr = arr + cont
                out[f"{name}_absorption"] = arr.reshape(shape + (size,))
            else:
                arr = np.zeros((t.size, len(MECHANISMS), size))
                if beta is not None:
                    arr[:, 0, :] = beta
                if cont is not None:
                    arr[:, 1, :] = cont
                out[f"{name}_absorption"] = arr.reshape(shape + (len(MECHANISMS), size))
        if output_format == "all":
            out["mechanism"] = list(MECHANISMS)
        return out

    def to_dataset(self, result, dims=None):
        """The reference's output ``Dataset`` (pyLBL/spectroscopy.py:208-236); needs xarray."""
        from xarray import DataArray, Dataset
        dims = list(dims) if dims is not None else [f"dim_{i}" for i in range(self.temperature.ndim)]
        units = {"units": "m-1"}
        data_vars = {"wavenumber": DataArray(self.grid, dims=("wavenumber",), attrs={"units": "cm-1"})}
        for name, value in result.items():
            if name == "wavenumber":
                continue
            if name == "mechanism":
                data_vars[name] = DataArray(value, dims=("mechanism",))
            elif value.ndim == len(dims) + 2:
                data_vars[name] = DataArray(value, dims=dims + ["mechanism", "wavenumber"], attrs=units)
            else:
                data_vars[name] = DataArray(value, dims=dims + ["wavenumber"], attrs=units)
        return Dataset(data_vars=data_vars)
