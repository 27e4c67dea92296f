"""Host-side logic of the Spectroscopy adapter that needs no GPU (the numbers are checked on
the GPU in tests/test_gpu_parity.py::test_spectroscopy_adapter_matches_the_reference_driver_loop)."""
import numpy as np
import pytest

from pylbl_b200 import Spectroscopy, number_density
from pylbl_b200.spectroscopy import MECHANISMS


class Member(object):
    def __init__(self, data):
        self.data = data


def atmosphere(shape=(2, 3)):
    t = np.full(shape, 250.)
    p = np.full(shape, 5.0e4)
    return {"temperature": t, "pressure": p, "gases": {"H2O": np.full(shape, 1e-3)}}


def test_accepts_dicts_and_data_members():
    grid = np.arange(1., 11., 0.01)
    a = atmosphere()
    s1 = Spectroscopy(a, grid, "no-such.db")
    wrapped = type("A", (), {})()
    wrapped.temperature = Member(a["temperature"])
    wrapped.pressure = Member(a["pressure"])
    wrapped.gases = {"H2O": Member(a["gases"]["H2O"])}
    s2 = Spectroscopy(wrapped, grid, type("Db", (), {"path": "no-such.db"})())
    assert s1.database == s2.database == "no-such.db"
    assert np.array_equal(s1.temperature, s2.temperature) and s1.temperature.dtype == np.float64
    assert list(s1.gases) == ["H2O"] and s2.gases["H2O"].shape == (2, 3)


def test_rejects_mismatched_shapes_and_formats():
    grid = np.arange(1., 11., 0.01)
    a = atmosphere()
    a["gases"]["CO2"] = np.full((3, 2), 4e-4)
    with pytest.raises(ValueError):
        Spectroscopy(a, grid, "no-such.db")
    s = Spectroscopy(atmosphere(), grid, "no-such.db")
    with pytest.raises(ValueError):
        s.compute_absorption(output_format="everything")


def test_mechanism_axis_and_number_density_follow_the_reference():
    # pyLBL/spectroscopy.py:126 and :18-29 (n = p*x/(kB*T), kB = 1.38064852e-23)
    assert MECHANISMS == ["lines", "continuum", "cross_section"]
    n = number_density(288.99, 98388., 6.637074e-3)
    assert n == pytest.approx(98388. * 6.637074e-3 / (1.38064852e-23 * 288.99), rel=1e-15)


def test_only_a_missing_molecule_is_tolerated(small_db):
    """pyLBL/spectroscopy.py:53-57 sets gas = None for molecules the database has no data for
    -- and for nothing else.  A database that cannot be opened must raise, not come back as an
    all-zero spectrum."""
    grid = np.arange(1., 11., 0.01)
    a = atmosphere()
    a["gases"]["XeF6"] = np.full((2, 3), 1e-9)
    s = Spectroscopy(a, grid, small_db)
    assert s._gas("XeF6") is None                      # alias not in the database
    bad = Spectroscopy(atmosphere(), grid, "/nonexistent/dir/file.db")
    with pytest.raises(ValueError):
        bad._gas("H2O")
    for fmt in ("gas", "total"):
        with pytest.raises(ValueError):
            bad.compute_absorption(fmt)


def test_band_edges_partition_the_grid(small_db):
    """Band sharding: contiguous cell ranges that cover the grid once, balanced
    by a cost model of the lines in and next to each cell.  Needs no GPU work -- but a handle,
    so it is skipped where none can be opened."""
    from pylbl_b200 import Gas, _lib
    if _lib.device_count() == 0:
        pytest.skip("a handle needs a CUDA device")
    gas = Gas(small_db, "CO2")
    for n_bands in (1, 2, 3, 8):
        edges = gas.band_edges((1, 2001, 100), n_bands)
        assert edges[0] == 0 and edges[-1] == 2000 and np.all(np.diff(edges) >= 0)


def test_to_dataset_layout_with_a_stand_in_for_xarray(monkeypatch):
    """`to_dataset` against the reference's `_create_output_dataset` (pyLBL/spectroscopy.py:208-236):
    which variables exist, their dimensions and units, for the three output formats.  xarray is
    not in this image; a stand-in that records what it is given is enough to check the layout
    (tests/test_pylbl_integration.py runs the real one where it is installed)."""
    import sys
    import types

    class DataArray(object):
        def __init__(self, data, dims=None, attrs=None):
            self.data, self.dims, self.attrs = np.asarray(data), tuple(dims), dict(attrs or {})

    class Dataset(object):
        def __init__(self, data_vars=None):
            self.data_vars = dict(data_vars)

    monkeypatch.setitem(sys.modules, "xarray", types.SimpleNamespace(DataArray=DataArray, Dataset=Dataset))
    grid = np.arange(1., 11., 0.01)
    s = Spectroscopy(atmosphere((2, 3)), grid, "no-such.db")
    n = grid.size
    everything = {"H2O_absorption": np.zeros((2, 3, len(MECHANISMS), n)), "mechanism": list(MECHANISMS)}
    ds = s.to_dataset(everything, dims=["y", "x"]).data_vars
    assert set(ds) == {"wavenumber", "mechanism", "H2O_absorption"}
    assert ds["wavenumber"].dims == ("wavenumber",) and ds["wavenumber"].attrs == {"units": "cm-1"}
    assert np.array_equal(ds["wavenumber"].data, grid)
    assert ds["mechanism"].dims == ("mechanism",) and list(ds["mechanism"].data) == list(MECHANISMS)
    assert ds["H2O_absorption"].dims == ("y", "x", "mechanism", "wavenumber")
    assert ds["H2O_absorption"].attrs == {"units": "m-1"}
    per_gas = s.to_dataset({"H2O_absorption": np.zeros((2, 3, n))}, dims=["y", "x"]).data_vars
    assert set(per_gas) == {"wavenumber", "H2O_absorption"}
    assert per_gas["H2O_absorption"].dims == ("y", "x", "wavenumber")       # the mechanism axis summed away
    total = s.to_dataset({"absorption": np.zeros((2, 3, n))}).data_vars
    assert total["absorption"].dims == ("dim_0", "dim_1", "wavenumber")
    assert total["absorption"].attrs == {"units": "m-1"}


class FillingRecorder(object):
    """Stands in for the library: records the calls; a submitted gas 'returns' a spectrum equal
    to its own volume mixing ratio in every point of every layer (so that what the adapter does
    with it afterwards can be checked); opening 'XeF6' fails the way a missing molecule does."""

    def __init__(self):
        self.calls = []

    def lbl_last_error(self):
        return b"Error: molecule XeF6 not found in database."

    def __getattr__(self, name):
        def call(*args):
            self.calls.append((name, args))
            if name == "lbl_gas_open" and args[1] == b"XeF6":
                raise ValueError("Error inside c functions.")
            if name == "lbl_gas_submit":
                ptr, n_layers, p, t, x, v0, vn, npv, cut, ped, prec, dst = args
                import ctypes
                n = (vn - v0) * npv
                k = np.ctypeslib.as_array(ctypes.cast(dst, ctypes.POINTER(ctypes.c_double)), shape=(n_layers, n))
                k[:] = np.asarray(x)[:, None]
            return 0
        return call


def test_per_gas_formats_against_the_reference_loop(monkeypatch):
    """`compute_absorption("gas" | "all")` with the library replaced: beta = n * k[:grid.size] per
    gas (pyLBL/spectroscopy.py:181-191), the mechanism axis of "all" with only "lines" filled
    (:131,225-234), zeros for a molecule the database does not hold (:53-57), every gas submitted
    before any is waited for, and remove_pedestal following the continuum backend (:163-164)."""
    from pylbl_b200 import gas_optics, spectroscopy
    rec = FillingRecorder()
    for module in (gas_optics, spectroscopy):
        monkeypatch.setattr(module._lib, "library", lambda: rec)
    monkeypatch.setattr(Spectroscopy, "PINNED_LIMIT", 0)            # ordinary staging arrays
    shape = (2, 3)
    a = atmosphere(shape)
    a["gases"]["CO2"] = np.full(shape, 4e-4)
    a["gases"]["XeF6"] = np.full(shape, 1e-12)
    grid = np.arange(1., 11., 0.01)[:-7]                            # the grid may stop short of vn
    s = Spectroscopy(a, grid, "spectral.db", continua_backend=None)
    per_gas = s.compute_absorption("gas")
    assert set(per_gas) == {"wavenumber", "H2O_absorption", "CO2_absorption", "XeF6_absorption"}
    n = number_density(a["temperature"], a["pressure"], a["gases"]["H2O"])
    assert per_gas["H2O_absorption"].shape == shape + (grid.size,)
    assert np.allclose(per_gas["H2O_absorption"], (n * 1e-3)[..., None])      # n * k, k == vmr here
    assert not per_gas["XeF6_absorption"].any()
    names = [c[0] for c in rec.calls if c[0] in ("lbl_gas_submit", "lbl_gas_wait")]
    assert names == ["lbl_gas_submit"] * 2 + ["lbl_gas_wait"] * 2             # all in flight, then collected
    assert all(c[1][9] == 0 for c in rec.calls if c[0] == "lbl_gas_submit")   # no continuum: no pedestal removal
    everything = s.compute_absorption("all", remove_pedestal=True)
    assert everything["mechanism"] == list(MECHANISMS)
    h2o = everything["H2O_absorption"]
    assert h2o.shape == shape + (len(MECHANISMS), grid.size)
    assert np.array_equal(h2o[..., 0, :], per_gas["H2O_absorption"]) and not h2o[..., 1:, :].any()
    assert [c[1][9] for c in rec.calls if c[0] == "lbl_gas_submit"][-2:] == [1, 1]
    s.close()
