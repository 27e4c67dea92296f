"""End-to-end against a real ``pyLBL.Spectroscopy`` with ``lines_backend="b200"`` (SURVEY.md
8(b), last row) -- for environments that have pyLBL and its dependencies installed; the build
image has neither xarray nor SQLAlchemy, so these are skipped there.  Shaped after the
reference's tests/test_spectroscopy.py:15-25 and tests/conftest.py:80-101, on a synthetic
database (the HITRAN download those use is not reproducible offline).

``Spectroscopy.to_dataset`` (the reference's output layout, pyLBL/spectroscopy.py:208-236) only
needs xarray and is tested on its own.
"""
import numpy as np
import pytest

from pylbl_b200 import synth

pytestmark = pytest.mark.gpu


def _dataset(xarray, atm, gases):
    """tests/conftest.py:80-101 of the reference: one `layer` dimension, CF standard names."""
    def variable(data, units, standard_name):
        return (["layer", ], data, {"units": units, "standard_name": standard_name})
    names = {"H2O": "water_vapor", "CO2": "carbon_dioxide", "O3": "ozone"}
    data_vars = {"pressure": variable(atm.p, "Pa", "air_pressure"),
                 "temperature": variable(atm.t, "K", "air_temperature")}
    for g in gases:
        data_vars[names[g]] = variable(atm.vmr[g], "mol mol-1", f"mole_fraction_of_{names[g]}_in_air")
    return xarray.Dataset(data_vars=data_vars)


def test_unmodified_pylbl_driver_with_the_b200_backend(small_db):
    xarray = pytest.importorskip("xarray")
    pyLBL = pytest.importorskip("pyLBL")
    import pylbl_b200
    from oracle import ReferenceGas
    assert pylbl_b200.registered and "b200" in pyLBL.plugins.molecular_lines

    class Db(object):            # the lines backends read only `.path`; no arts-crossfit data
        path = small_db

        def arts_crossfit(self, name):
            raise pyLBL.database.CrossSectionNotFoundError(name)

    atm = synth.fixture_atmosphere()
    gases = ["H2O", "CO2", "O3"]
    grid = np.arange(1., 1200., 0.1)
    kwargs = dict(continua_backend="mt_ckd", cross_sections_backend="arts_crossfit")
    ours = pyLBL.Spectroscopy(_dataset(xarray, atm, gases), grid, Db(), lines_backend="b200", **kwargs)
    beta = ours.compute_absorption(output_format="all")
    theirs = pyLBL.Spectroscopy(_dataset(xarray, atm, gases), grid, Db(), lines_backend="pyLBL", **kwargs)
    want = theirs.compute_absorption(output_format="all")
    for g in gases:
        a, b = beta[f"{g}_absorption"].data, want[f"{g}_absorption"].data
        assert a.shape == b.shape == (4, 3, grid.size)
        scale = np.abs(b[:, 0, :]).max(axis=1, keepdims=True)
        assert np.abs(a[:, 0, :] - b[:, 0, :]).max() <= 1e-9 * scale.max()
        assert np.array_equal(a[:, 1:, :], b[:, 1:, :])        # the other mechanisms are untouched


def test_to_dataset_has_the_references_layout(small_db):
    xarray = pytest.importorskip("xarray")
    from pylbl_b200 import Spectroscopy
    atm = synth.fixture_atmosphere()
    shape = (2, 2)
    state = {"temperature": atm.t.reshape(shape), "pressure": atm.p.reshape(shape),
             "gases": {g: atm.vmr[g].reshape(shape) for g in ("H2O", "CO2")}}
    grid = synth.grid_from_bounds(1, 301, 10)
    s = Spectroscopy(state, grid, small_db)
    ds = s.to_dataset(s.compute_absorption("all"), dims=["y", "x"])
    assert isinstance(ds, xarray.Dataset)
    assert ds["H2O_absorption"].dims == ("y", "x", "mechanism", "wavenumber")      # spectroscopy.py:131
    assert ds["H2O_absorption"].attrs["units"] == "m-1" and ds["wavenumber"].attrs["units"] == "cm-1"
    assert list(ds["mechanism"].data) == ["lines", "continuum", "cross_section"]
    total = s.to_dataset(s.compute_absorption("total"), dims=["y", "x"])
    assert total["absorption"].dims == ("y", "x", "wavenumber")
    s.close()
