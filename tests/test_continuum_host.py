"""Host-side logic of `Continuum` that needs no GPU: what the library is handed (the state the
band formulas read of the reference's vmr dictionary, the names, the layer tiles), and the
errors the reference would raise.  The library is replaced by a recorder; the numbers are
checked on the GPU (tests/test_gpu_parity.py::test_continuum_against_the_oracle)."""
import numpy as np
import pytest

from pylbl_b200 import Continuum, continua_of, continuum


class Recorder(object):
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def call(*args):
            self.calls.append((name, args))
            return 0
        return call


@pytest.fixture
def cont(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(continuum._lib, "library", lambda: rec)
    return rec, Continuum(device=0)


def test_table_is_loaded_once_spectrum_by_spectrum(cont):
    rec, c = cont
    names = [n for n, _ in rec.calls]
    assert names[0] == "lbl_continuum_create" and names[-1] == "lbl_continuum_finalize"
    spectra = [a for n, a in rec.calls if n == "lbl_continuum_set_spectrum"]
    assert len(spectra) >= 10                                   # the reference file holds 17 arrays
    for ptr, key, lower, upper, resolution, size, data in spectra:
        # the grid of a spectrum: lower + i*resolution, the last point on the upper bound (mt_ckd/utils.py:128-142)
        assert size == data.size and resolution > 0.
        assert abs(lower + (size - 1) * resolution - upper) <= 1e-6 * max(1., abs(upper))


def test_continua_of_a_gas_follow_the_reference():
    # the entry points of setup.py:45-57 and spectroscopy.py:194-198: water vapour has two continua
    assert continua_of("H2O") == ["H2OForeign", "H2OSelf"]
    for f in ("CO2", "O3", "N2", "O2"):
        assert continua_of(f) == [f]
    assert continua_of("CH4") == [] and continua_of("CO") == []


def test_state_columns_and_layer_tiles(cont, monkeypatch):
    rec, c = cont
    n = 7
    t, p = np.linspace(200., 300., n), np.linspace(1e3, 1e5, n)
    vmr = {"H2O": np.linspace(1e-6, 1e-2, n), "CO2": np.full(n, 4e-4), "O3": np.full(n, 1e-6),
           "N2": np.full(n, 0.78), "O2": np.full(n, 0.21), "CH4": np.full(n, 2e-6)}
    monkeypatch.setattr(Continuum, "MAX_LAYERS", 3)
    out = c.spectra(["H2OForeign", "H2OSelf", "CO2"], t, p, vmr, bounds=(1, 11, 10))
    assert out.shape == (7, 100)
    calls = [a for nme, a in rec.calls if nme == "lbl_continuum_compute"]
    assert [a[2] for a in calls] == [3, 3, 1]                   # layers per tile
    assert all(a[1] == b"H2OForeign,H2OSelf,CO2" for a in calls)
    state = np.concatenate([a[5] for a in calls])
    assert state.shape == (7, 6)
    for k, name in enumerate(("H2O", "CO2", "O3", "N2", "O2")):
        assert np.array_equal(state[:, k], vmr[name])
    assert np.allclose(state[:, 5], sum(vmr.values()))          # every gas of the atmosphere counts
    assert np.array_equal(np.concatenate([a[3] for a in calls]), t)
    # into an accumulator: rows offset by row0, no host array
    rec.calls.clear()
    assert c.spectra("N2", t, p, vmr, bounds=(1, 11, 10), mix=object(), row0=5) is None
    calls = [a for nme, a in rec.calls if nme == "lbl_continuum_compute"]
    assert [a[10] for a in calls] == [5, 8, 11] and all(a[11] is None for a in calls)


def test_missing_gases_raise_like_the_reference(cont):
    rec, c = cont
    t, p = np.array([250.]), np.array([5e4])
    with pytest.raises(KeyError):                               # dry-air density needs H2O (mt_ckd/utils.py:44)
        c.spectra("CO2", t, p, {"CO2": np.array([4e-4])}, bounds=(1, 11, 10))
    with pytest.raises(KeyError):                               # nitrogen.py reads O2 too
        c.spectra("N2", t, p, {"H2O": np.array([1e-3]), "N2": np.array([0.78])}, bounds=(1, 11, 10))
    # a gas no requested formula needs may be absent: its column is NaN, never read
    c.spectra("H2OSelf", t, p, {"H2O": np.array([1e-3])}, bounds=(1, 11, 10))
    state = [a for n, a in rec.calls if n == "lbl_continuum_compute"][-1][5]
    assert state[0, 0] == 1e-3 and np.isnan(state[0, 1]) and state[0, 5] == 1e-3
    with pytest.raises(ValueError):
        c.spectra("H2OSelf", t, p, {"H2O": np.array([1e-3])}, bounds=(1, 11, 10), out=np.empty((1, 99)))
