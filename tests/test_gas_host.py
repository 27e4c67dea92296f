"""Host-side logic of `Gas` that needs no GPU: how a call is cut over several devices (layers or
spectral bands), what each piece hands the library, the prefetch cache behind the scalar plugin
call, and argument checks.  The library is replaced by a recorder; the numbers are checked on the
GPU (tests/test_gpu_parity.py)."""
import ctypes

import numpy as np
import pytest

from pylbl_b200 import Gas, gas_optics, synth


class Recorder(object):
    def __init__(self, band_edges=None):
        self.calls = []
        self.band_edges = band_edges

    def __getattr__(self, name):
        def call(*args):
            self.calls.append((name, args))
            if name == "lbl_gas_band_edges" and self.band_edges is not None:
                args[-1][:] = self.band_edges
            return 0
        return call

    def named(self, name):
        return [a for n, a in self.calls if n == name]


@pytest.fixture
def rec(monkeypatch):
    r = Recorder()
    monkeypatch.setattr(gas_optics._lib, "library", lambda: r)
    return r


def address(pointer):
    return ctypes.cast(pointer, ctypes.c_void_p).value


def column(n):
    return np.linspace(220., 290., n), np.linspace(1e3, 1e5, n), np.linspace(1e-6, 1e-2, n)


def test_layers_are_cut_into_contiguous_shards_one_per_device(rec):
    gas = Gas("spectral.db", "H2O", devices=[0, 1, 2])
    t, p, x = column(10)
    out = np.empty((10, 1000))
    k = gas.absorption_coefficients(t, p, x, bounds=(1, 11, 100), remove_pedestal=True, cut_off=25, out=out)
    assert k is out
    opens = rec.named("lbl_gas_open")
    assert sorted(a[2] for a in opens) == [0, 1, 2] and all(a[0] == b"spectral.db" and a[1] == b"H2O" for a in opens)
    calls = sorted(rec.named("lbl_gas_compute"), key=lambda a: a[3][0])      # by first temperature
    edges = [0, 3, 6, 10]
    assert len(calls) == 3
    for d, c in enumerate(calls):
        ptr, n_layers, pressure, temperature, vmr, v0, vn, npv, cut, ped, prec, dst = c
        lo, hi = edges[d], edges[d + 1]
        assert n_layers == hi - lo and (v0, vn, npv, cut, ped) == (1, 11, 100, 25, 1)
        # the reference's argument order: pressure before temperature (c_lib/absorption.c:19-30)
        assert np.array_equal(pressure, p[lo:hi]) and np.array_equal(temperature, t[lo:hi])
        assert np.array_equal(vmr, x[lo:hi])
        assert address(dst) == out.ctypes.data + lo * out.strides[0]       # each shard writes its own rows
    assert len(rec.named("lbl_gas_wait")) == 3
    # fewer layers than devices: only as many shards as layers
    rec.calls.clear()
    gas.absorption_coefficients(t[:2], p[:2], x[:2], bounds=(1, 11, 100))
    assert [a[1] for a in rec.named("lbl_gas_compute")] == [1, 1]


def test_bands_one_per_device_write_their_columns(monkeypatch):
    r = Recorder(band_edges=[0, 4, 4, 10])                      # the middle band is empty
    monkeypatch.setattr(gas_optics._lib, "library", lambda: r)
    gas = Gas("spectral.db", "CO2", devices=[0, 1, 2], shard="band")
    t, p, x = column(5)
    out = np.empty((5, 1000))
    gas.absorption_coefficients(t, p, x, bounds=(1, 11, 100), out=out)
    asked = r.named("lbl_gas_band_edges")
    assert len(asked) == 1 and asked[0][1:6] == (1, 11, 100, 25, 3)
    calls = sorted(r.named("lbl_gas_submit_band"), key=lambda a: a[11])
    assert [(a[11], a[12]) for a in calls] == [(0, 4), (4, 10)]              # no call for the empty band
    for a in calls:
        assert a[1] == 5 and np.array_equal(a[3], t)                         # every band: all layers
        assert address(a[13]) == out.ctypes.data + a[11] * 100 * 8           # its first column
        assert a[14] == 1000                                                 # pitch: a whole-grid row
    with pytest.raises(ValueError):
        Gas("spectral.db", "CO2", shard="columns")


def test_band_destination_must_have_contiguous_rows(rec):
    gas = Gas("spectral.db", "CO2", devices=[0])
    t, p, x = column(3)
    wide = np.empty((3, 1000))
    gas.absorption_band(t, p, x, (1, 11, 100), (2, 5), out=wide[:, 200:500])    # a column slice is fine
    assert rec.named("lbl_gas_submit_band")[-1][14] == 1000
    dense = gas.absorption_band(t, p, x, (1, 11, 100), (2, 5))
    assert dense.shape == (3, 300) and rec.named("lbl_gas_submit_band")[-1][14] == 300
    for bad in (np.empty((3, 299)), np.empty((3, 600))[:, ::2], np.empty((3, 300), dtype=np.float32)):
        with pytest.raises(ValueError):
            gas.absorption_band(t, p, x, (1, 11, 100), (2, 5), out=bad)


def test_prefetched_rows_are_handed_out_once(rec):
    """What lets an unmodified pyLBL driver loop (pyLBL/spectroscopy.py:179-191) run a column as one
    batch: `prefetch` keeps the rows by (T, p, x, grid, cut_off, remove_pedestal); the scalar calls
    that follow take them without touching the library, each row once."""
    gas = Gas("spectral.db", "O3", devices=[0])
    grid = synth.grid_from_bounds(1, 11, 100)
    t, p, x = column(4)
    gas.prefetch(t, p, x, grid, remove_pedestal=True)
    assert len(rec.named("lbl_gas_compute")) == 1 and rec.named("lbl_gas_compute")[0][1] == 4
    rec.calls.clear()
    rows = [gas.absorption_coefficient(t[i], p[i], x[i], grid, remove_pedestal=True) for i in range(4)]
    assert rec.named("lbl_gas_compute") == [] and all(r.shape == (1000,) for r in rows)
    # a second request for the same state, or the same state with other options, computes
    gas.absorption_coefficient(t[0], p[0], x[0], grid, remove_pedestal=True)
    assert len(rec.named("lbl_gas_compute")) == 1 and rec.named("lbl_gas_compute")[0][1] == 1
    gas.prefetch(t, p, x, grid, remove_pedestal=True)
    rec.calls.clear()
    gas.absorption_coefficient(t[1], p[1], x[1], grid, remove_pedestal=False)
    gas.absorption_coefficient(t[1], p[1], x[1], grid, remove_pedestal=True, cut_off=10)
    assert len(rec.named("lbl_gas_compute")) == 2


def test_submit_and_wait_use_the_non_blocking_entry_point(rec):
    gas = Gas("spectral.db", "CO", devices=[1])
    t, p, x = column(6)
    assert gas.wait() is None                                   # nothing in flight
    gas.submit(t, p, x, bounds=(1, 11, 100), remove_pedestal=True)
    assert [n for n, _ in rec.calls if n.startswith("lbl_gas_")][-1] == "lbl_gas_submit"
    call = rec.named("lbl_gas_submit")[0]
    assert call[1] == 6 and call[11] is None                    # spectra stay on the device
    assert rec.named("lbl_gas_wait") == []
    gas.wait()
    assert len(rec.named("lbl_gas_wait")) == 1
    with pytest.raises(ValueError):
        gas.submit(t, p, x, bounds=(1, 11, 100), out=np.empty((6, 999)))
