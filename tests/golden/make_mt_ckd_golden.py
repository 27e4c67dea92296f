"""Generates tests/golden/mt_ckd_reference.npz by running the REFERENCE's own continuum code
(/root/reference/pyLBL/mt_ckd/*.py, unmodified) in this container.

The reference reads its coefficient file through ``netCDF4.Dataset`` (mt_ckd/utils.py:3,131-136),
which is not installed here; this script registers a stand-in ``netCDF4`` module whose
``Dataset`` serves ``variables[name][:]`` and ``getncattr`` from tools/hdf5_min.py (a reader for
the HDF5 subset that file uses), and a bare ``pyLBL`` package object so that ``pyLBL.mt_ckd``
imports without pyLBL/__init__.py (which needs xarray).  Nothing of the reference is copied.

It first repeats the reference's own known-answer test (tests/test_mt_ckd.py:15-46: per-band
sums on the surface layer of the fixture atmosphere) -- which pins the HDF5 reader and this
environment against numbers the reference's authors recorded -- and then stores
``BandedContinuum.spectra`` (utils.py:157-174) for every continuum on three grids and two layers.

    python tests/golden/make_mt_ckd_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import hdf5_min  # noqa: E402

REFERENCE = Path("/root/reference")


class _Variable(object):
    def __init__(self, dataset):
        self._d = dataset

    def __getitem__(self, key):
        return self._d.data[key]

    def getncattr(self, name):
        return float(np.asarray(self._d.attrs[name]).ravel()[0])


class _Dataset(object):
    _cache = {}

    def __init__(self, path, mode="r"):
        if path not in self._cache:
            self._cache[path] = hdf5_min.read_file(path)
        self.variables = {k: _Variable(v) for k, v in self._cache[path].items() if v.data is not None}

    def __enter__(self):
        return self

    def __exit__(self, *args):
        return False


def import_reference_mt_ckd():
    stub = types.ModuleType("netCDF4")
    stub.Dataset = _Dataset
    sys.modules["netCDF4"] = stub
    package = types.ModuleType("pyLBL")
    package.__path__ = [str(REFERENCE / "pyLBL")]
    sys.modules["pyLBL"] = package
    from pyLBL.mt_ckd import carbon_dioxide, nitrogen, oxygen, ozone, water_vapor
    return {
        "CO2": carbon_dioxide.CarbonDioxideContinuum,
        "H2OForeign": water_vapor.WaterVaporForeignContinuum,
        "H2OSelf": water_vapor.WaterVaporSelfContinuum,
        "N2": nitrogen.NitrogenContinuum,
        "O2": oxygen.OxygenContinuum,
        "O3": ozone.OzoneContinuum,
    }


# the reference's fixture atmosphere, tests/conftest.py:54-78
PRESSURE = np.asarray([117., 1032., 11419., 98388.])
TEMPERATURE = np.asarray([269.01, 227.74, 203.37, 288.99])
VMR = {
    "H2O": np.asarray([5.244536e-06, 4.763972e-06, 3.039952e-06, 6.637074e-03]),
    "CO2": np.asarray([0.00036, 0.00036, 0.00036, 0.00035999]),
    "O3": np.asarray([2.936688e-06, 7.415223e-06, 2.609510e-07, 6.859128e-08]),
    "N2O": np.asarray([1.050928e-08, 1.319584e-07, 2.895416e-07, 3.199949e-07]),
    "CH4": np.asarray([2.947482e-07, 8.817705e-07, 1.588336e-06, 1.700002e-06]),
    "CO": np.asarray([3.621464e-08, 1.761450e-08, 3.315927e-08, 1.482969e-07]),
    "O2": np.asarray([0.209, 0.209, 0.2090003, 0.208996]),
    "N2": np.asarray([0.78, 0.78, 0.78, 0.78]),
}
# tests/test_mt_ckd.py:15-27
KNOWN_ANSWERS = {
    "CO2": [21.284607102488753, ],
    "H2OForeign": [131.87162317621952, ],
    "H2OSelf": [13.482864611247933, ],
    "N2": [0.7612890022253513, 0.5875825355004741, 0.00414557543788256, ],
    "O2": [0.24690308716508605, 0.11052072297118236, 0.03200556021322852,
           0.04514938962400228, 0.03897535512343981, 285.7607588975901,
           4419601.794329887, ],
    "O3": [0.0006562127133778276, 1.7334221226752753, 0.05197265302394795, ],
}
GRIDS = {
    "coarse": np.arange(1., 5000., 2.),            # the infrared at large
    "bandhead": np.arange(2380., 2440., 0.02),     # CO2 band head: the tdep/x-factor sub-grids
    "uv": np.arange(7000., 60000., 50.),           # O2 and O3 near-infrared to ultra-violet bands
}
LAYERS = [1, 3]


def main():
    classes = import_reference_mt_ckd()
    continua = {name: cls() for name, cls in classes.items()}
    out = {"layers": np.asarray(LAYERS), "pressure": PRESSURE, "temperature": TEMPERATURE}
    for key, value in VMR.items():
        out["vmr_" + key] = value
    # (1) the reference's known-answer test, as it calls it: band.spectra(T, p[Pa], vmr)
    index = -1
    vmr = {k: v[index] for k, v in VMR.items()}
    worst = 0.
    for name, continuum in continua.items():
        sums = []
        for band, want in zip(continuum.bands, KNOWN_ANSWERS[name]):
            got = float(np.sum(band.spectra(TEMPERATURE[index], PRESSURE[index], vmr)))
            worst = max(worst, abs(got - want) / abs(want))
            sums.append(got)
        out["band_sums_" + name] = np.asarray(sums)
        assert len(continuum.bands) == len(KNOWN_ANSWERS[name])
    print(f"reference known answers reproduced to {worst:.2e} (pytest.approx: 1e-6)")
    assert worst <= 1e-6
    # (2) the interpolated continua the driver adds (spectroscopy.py:194-198), per grid and layer
    for gname, grid in GRIDS.items():
        out["grid_" + gname] = grid
        for name, continuum in continua.items():
            rows = []
            for layer in LAYERS:
                vmr = {k: v[layer] for k, v in VMR.items()}
                rows.append(continuum.spectra(TEMPERATURE[layer], PRESSURE[layer], vmr, grid))
            out[f"{name}_{gname}"] = np.asarray(rows)
    path = Path(__file__).resolve().parent / "mt_ckd_reference.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
