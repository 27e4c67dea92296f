"""Host-side logic of `Mixture.total_absorption` that needs no GPU: the order of the calls it
makes for layer groups, what each call is given, and how it adapts the group count to the copy
tail the library reports.  The library is replaced by a recorder; the numbers are checked on
the GPU (tests/test_gpu_parity.py::test_mixture_total_absorption and the tests next to it)."""
import ctypes

import numpy as np
import pytest

from pylbl_b200 import Mixture, mixture, number_density


class Recorder(object):
    """Stands in for the ctypes library: records every call, returns 0."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def call(*args):
            self.calls.append((name, args))
            return 0
        return call


class Handle(object):
    def __init__(self, ptr, n_lines, recorder, tail_ms=0.):
        self.ptr, self.n_lines, self.recorder, self.tail_ms = ptr, n_lines, recorder, tail_ms

    def stats(self):
        zero = {k: 0 for k in ("evals", "executed", "h2d_bytes", "d2h_bytes", "sum_launches", "total_launches")}
        zero.update({k: 0. for k in ("scale_ms", "sum_ms", "fixup_ms", "pedestal_ms")})
        # what the last submit on this handle covered
        layers = [a[1] for n, a in self.recorder.calls if n == "lbl_gas_submit_mix" and a[0] == self.ptr]
        zero.update(n_lines=self.n_lines, n_layers=layers[-1] if layers else 0, total_ms=1.0,
                    evals=layers[-1] if layers else 0, copy_tail_ms=self.tail_ms)
        return zero


class FakeGas(object):
    precision = 0

    def __init__(self, handle):
        self.handle = handle
        self.last_stats = None

    def _handle(self, device):
        return self.handle


@pytest.fixture
def setup(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(mixture._lib, "library", lambda: rec)
    gases = {"CO2": FakeGas(Handle(103, 90000, rec)), "CO": FakeGas(Handle(101, 2000, rec)),
             "H2O": FakeGas(Handle(102, 70000, rec))}
    mix = Mixture.from_gases(gases, device=0)
    return rec, gases, mix


def column(n_layers):
    t = np.linspace(220., 290., n_layers)
    p = np.linspace(1.0e3, 1.0e5, n_layers)
    vmr = {"CO2": np.full(n_layers, 4e-4), "CO": np.full(n_layers, 1e-7), "H2O": np.linspace(1e-6, 1e-2, n_layers)}
    return t, p, vmr


def submits(rec):
    return [a for n, a in rec.calls if n == "lbl_gas_submit_mix"]


def test_one_group_submits_every_gas_once_fewest_lines_first(setup):
    rec, gases, mix = setup
    t, p, vmr = column(10)
    out = mix.total_absorption(t, p, vmr, bounds=(1, 11, 100))
    assert out.shape == (10, 1000)
    calls = submits(rec)
    assert [c[0] for c in calls] == [101, 102, 103]            # CO, H2O, CO2: the longest list last
    for c in calls:
        ptr, n_layers, pressure, temperature, x, v0, vn, npv, cut, ped, prec, mixh, row0, scale, host = c
        assert (n_layers, v0, vn, npv, cut, ped, row0) == (10, 1, 11, 100, 25, 1, 0)
        assert np.array_equal(pressure, p) and np.array_equal(temperature, t)
    # beta = n * k: the scale handed over is the number density of that gas (spectroscopy.py:18-29)
    assert np.allclose(calls[0][13], number_density(t, p, vmr["CO"]))
    assert np.allclose(calls[2][13], number_density(t, p, vmr["CO2"]))
    # only the last gas brings the host array along; the others add on the device and return
    assert calls[0][14] is None and calls[1][14] is None and calls[2][14] is not None
    names = [n for n, _ in rec.calls]
    assert names.index("lbl_mix_open") < names.index("lbl_gas_submit_mix") < names.index("lbl_mix_wait")
    assert mix.last_layer_groups == 1


def test_layer_groups_are_group_major_with_row_offsets(setup):
    rec, gases, mix = setup
    t, p, vmr = column(10)
    mix.total_absorption(t, p, vmr, bounds=(1, 11, 100), layer_groups=3)
    calls = submits(rec)
    assert [c[0] for c in calls] == [101, 102, 103] * 3
    edges = [0, 3, 6, 10]
    for g in range(3):
        for c in calls[3 * g:3 * g + 3]:
            assert c[1] == edges[g + 1] - edges[g] and c[12] == edges[g]        # layers, first row
            assert np.array_equal(c[3], t[edges[g]:edges[g + 1]])
            assert np.array_equal(c[13], number_density(t, p, vmr[{101: "CO", 102: "H2O", 103: "CO2"}[c[0]]])
                                  [edges[g]:edges[g + 1]])
    # the last gas of a group copies its rows out; in every group but the last as ONE piece
    groups = [a for n, a in rec.calls if n == "lbl_gas_set_copy_groups"]
    assert groups == [(103, 1), (103, 1), (103, 0)]
    # a handle is waited for before it is given its next group
    order = [(n, a[0]) for n, a in rec.calls if n in ("lbl_gas_submit_mix", "lbl_gas_wait")]
    for ptr in (101, 102, 103):
        mine = [n for n, q in order if q == ptr]
        assert mine[:5] == ["lbl_gas_submit_mix", "lbl_gas_wait", "lbl_gas_submit_mix", "lbl_gas_wait",
                            "lbl_gas_submit_mix"]
    # statistics add up over the groups
    assert gases["CO2"].last_stats[0]["n_layers"] == 10 and gases["CO2"].last_stats[0]["evals"] == 10
    assert mix.last_layer_groups == 3


def test_group_count_follows_the_reported_copy_tail(setup, monkeypatch):
    rec, gases, mix = setup
    t, p, vmr = column(64)
    bounds = (1, 11, 100)
    mix.total_absorption(t, p, vmr, bounds=bounds)
    assert mix.last_layer_groups == 1
    # nothing uncovered: stays at one group
    mix.total_absorption(t, p, vmr, bounds=bounds)
    assert mix.last_layer_groups == 1
    # the last gas reports a copy tail far above 12 % of the call: one more group each call, up to 4
    gases["CO2"].handle.tail_ms = 1.0e6
    seen = []
    for _ in range(6):
        mix.total_absorption(t, p, vmr, bounds=bounds)
        seen.append(mix.last_layer_groups)
    assert seen == [1, 2, 3, 4, 4, 4]
    # too few layers for that many groups: one group, whatever was learnt
    t8, p8, vmr8 = column(8)
    mix.total_absorption(t8, p8, vmr8, bounds=bounds)
    assert mix.last_layer_groups == 1
    # an explicit request overrides the automatic choice and teaches it nothing
    mix.total_absorption(t, p, vmr, bounds=bounds, layer_groups=2)
    assert mix.last_layer_groups == 2
    mix.total_absorption(t, p, vmr, bounds=bounds)
    assert mix.last_layer_groups == 4


def test_out_must_match(setup):
    rec, gases, mix = setup
    t, p, vmr = column(4)
    with pytest.raises(ValueError):
        mix.total_absorption(t, p, vmr, bounds=(1, 11, 100), out=np.empty((4, 999)))
    with pytest.raises(ValueError):
        mix.total_absorption(t, p, vmr, bounds=(1, 11, 100), out=np.empty((4, 1000), dtype=np.float32))
