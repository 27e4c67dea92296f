"""Kernel logic on the CPU: the per-thread bodies of K1-K4 (pylbl_b200/csrc/lbl_threads.cuh)
compiled as plain C++ (tests/emu/emu.cpp) and compared with the oracle.

This covers the host-visible logic of the gather formulation -- window membership per
point, the five-segment line ranges, the near/far split, the pedestal recurrence -- on
machines without a GPU.  It is not the product path (the reciprocal seed is emulated).
"""
import ctypes
import subprocess
from ctypes import POINTER, c_int, c_longlong
from pathlib import Path

import numpy as np
import pytest
from numpy.ctypeslib import ndpointer

from oracle import OracleGas
from pylbl_b200 import synth

from helpers import FP32_TOL, FP64_TOL, relative_error, scaled_error

EMU_DIR = Path(__file__).resolve().parent / "emu"


@pytest.fixture(scope="module")
def emu():
    so, src = EMU_DIR / "libemu.so", EMU_DIR / "emu.cpp"
    core = EMU_DIR.parent.parent / "pylbl_b200" / "csrc"
    newest = max(p.stat().st_mtime for p in (src, core / "lbl_core.cuh", core / "lbl_threads.cuh",
                                             core / "lbl_bands.h"))
    if not so.exists() or so.stat().st_mtime < newest:
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                        "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    f64 = ndpointer(np.float64, flags="C_CONTIGUOUS")
    i32 = ndpointer(np.int32, flags="C_CONTIGUOUS")
    lib.emu_absorption.argtypes = [c_int, f64, f64, f64, c_int, c_int, c_int, f64, c_int] + \
        [f64] * 7 + [i32, f64, c_int, c_int, f64, f64, c_int, c_int, c_int, c_int, POINTER(c_longlong)]
    lib.emu_absorption.restype = c_int
    lib.emu_partition_bands.argtypes = [f64, f64, c_int, c_int, i32]
    lib.emu_partition_bands.restype = c_int
    return lib


def run(emu, gas, t, p, x, bounds, ped, cut=25, points=None, fp32=False):
    d = gas.data
    v0, vn, npv = bounds
    n = (vn - v0) * npv
    t, p, x = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, p, x))
    k = np.zeros(t.size * n)
    if points is None:
        # the library's choice: cell-tiled far-field kernel (coded 0 here) on fine grids
        points = 0 if (npv >= 64 and not fp32) else [q for q in (5, 4, 8, 10, 2, 1) if npv % q == 0][0]
    evals = c_longlong(0)
    rc = emu.emu_absorption(t.size, p, t, x, v0, vn, npv, k, d["nu"].size, d["nu"], d["sw"],
                            d["gamma_air"], d["gamma_self"], d["n_air"], d["elower"],
                            d["delta_air"], d["local_iso_id"], d["mass"], d["num_iso"],
                            d["num_t"], d["tips_t"], d["tips_q"], cut, int(ped), points,
                            1 if fp32 else 0, ctypes.byref(evals))
    assert rc == 0
    return k.reshape(t.size, n), int(evals.value)


@pytest.mark.parametrize("bounds", [(1, 601, 10), (1, 301, 100), (1, 900, 1), (1, 500, 4),
                                    (1, 400, 7), (1, 201, 8), (100, 161, 1000), (1, 121, 64),
                                    (1, 81, 250), (1, 31, 2000)])
@pytest.mark.parametrize("ped", [0, 1])
def test_emulated_vs_oracle(emu, small_db, atmosphere, bounds, ped):
    for formula in ("H2O", "O3"):
        gas = OracleGas(small_db, formula)
        k, evals = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr[formula], bounds, ped)
        total = 0
        for layer in range(atmosphere.t.size):
            k_ref = gas.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr[formula][layer], *bounds, ped)
            total += gas.last_evals
            if not np.any(k_ref):
                assert not np.any(k[layer])
                continue
            assert scaled_error(k[layer], k_ref, bounds[2]) <= FP64_TOL
            if not ped:
                assert relative_error(k[layer], k_ref) <= FP64_TOL
        assert evals == total


@pytest.mark.parametrize("points", [1, 2, 5, 10])
def test_points_per_thread_variants_agree(emu, small_db, atmosphere, points):
    gas = OracleGas(small_db, "CO2")
    bounds = (600, 761, 10)
    # 600 - 26 > first row: early break -> zeros for every variant
    k, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"], bounds, 1, points=points)
    assert not np.any(k)
    bounds = (1, 761, 10)
    k, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"], bounds, 1, points=points)
    k_ref = gas.absorption(atmosphere.t[2], atmosphere.p[2], atmosphere.vmr["CO2"][2], *bounds, 1)
    assert scaled_error(k[2], k_ref, 10) <= FP64_TOL


@pytest.mark.parametrize("cut", [0, 1, 5, 40, 70])
def test_cut_off_variants(emu, small_db, atmosphere, cut):
    gas = OracleGas(small_db, "H2O")
    bounds = (1, 300, 10)
    for ped in (0, 1):
        k, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["H2O"], bounds, ped, cut=cut)
        for layer in (0, 3):
            k_ref = gas.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["H2O"][layer], *bounds, ped, cut)
            assert scaled_error(k[layer], k_ref, 10, max(cut, 1)) <= FP64_TOL


def test_unsorted_rows(emu, tmp_path, atmosphere):
    lines = synth.make_line_list("CO2", 400, 560.0, 760.0, seed=3)
    perm = np.random.default_rng(5).permutation(400)
    path = str(tmp_path / "unsorted.db")
    synth.write_database(path, {"CO2": {k: v[perm] for k, v in lines.items()}})
    gas = OracleGas(path, "CO2")
    bounds = (540, 781, 20)
    for ped in (0, 1):
        k, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"], bounds, ped)
        for layer in range(4):
            k_ref = gas.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["CO2"][layer], *bounds, ped)
            assert scaled_error(k[layer], k_ref, 20) <= FP64_TOL


def test_large_pressure_shift_near_integer_boundaries(emu, tmp_path):
    """Lines sitting on integer wavenumbers with large shifts: the window cell must come from
    the SHIFTED centre, bit-exactly (spectra.c:22,48)."""
    lines = synth.make_line_list("O2", 200, 30.0, 130.0, seed=9)
    lines["nu"] = np.sort(np.round(lines["nu"]) + np.tile([0.0, 1e-6, -1e-6, 0.5], 50))
    lines["delta_air"] = np.tile([-0.02, 0.02, 0.0199, -0.0003], 50)
    path = str(tmp_path / "shift.db")
    synth.write_database(path, {"O2": lines})
    gas = OracleGas(path, "O2")
    t = np.array([250.0, 296.0]); p = np.array([101325.0, 5.0e4]); x = np.array([0.209, 0.209])
    bounds = (5, 161, 20)
    for ped in (0, 1):
        k, evals = run(emu, gas, t, p, x, bounds, ped)
        total = 0
        for layer in range(2):
            k_ref = gas.absorption(t[layer], p[layer], x[layer], *bounds, ped)
            total += gas.last_evals
            assert scaled_error(k[layer], k_ref, 20) <= FP64_TOL
        assert evals == total


@pytest.mark.parametrize("bounds", [(1, 601, 10), (1, 301, 32), (1, 500, 4), (1, 41, 50)])
@pytest.mark.parametrize("ped", [0, 1])
def test_fp32_mode_within_stated_tolerance(emu, small_db, atmosphere, bounds, ped):
    """Opt-in FP32 far-wing arithmetic: 1e-4 of the local scale (and pointwise without the
    pedestal); windows and evaluation counts stay exact."""
    worst = 0.0
    for formula in ("H2O", "CO2"):
        gas = OracleGas(small_db, formula)
        k, evals = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr[formula], bounds, ped,
                       fp32=True)
        total = 0
        for layer in range(atmosphere.t.size):
            k_ref = gas.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr[formula][layer], *bounds, ped)
            total += gas.last_evals
            if not np.any(k_ref):
                assert not np.any(k[layer])
                continue
            err = scaled_error(k[layer], k_ref, bounds[2])
            worst = max(worst, err)
            assert err <= FP32_TOL
            if not ped:
                assert relative_error(k[layer], k_ref) <= FP32_TOL
        assert evals == total
    assert worst > 1e-12   # it really is the FP32 path


def test_far_field_kernel_against_direct_kernel(emu, small_db, atmosphere):
    """The cell-tiled kernel's polynomial far field reproduces the direct summation far below
    the parity tolerance (the interpolant of each far line is exact to ~1e-16 of the line)."""
    gas = OracleGas(small_db, "CO2")
    for bounds in ((1, 201, 100), (600, 640, 500)):
        for ped in (0, 1):
            fast, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"], bounds, ped, points=0)
            direct, _ = run(emu, gas, atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"], bounds, ped, points=5)
            for layer in range(4):
                if np.any(direct[layer]):
                    assert scaled_error(fast[layer], direct[layer], bounds[2]) <= 1e-11


@pytest.mark.parametrize("bounds,cut", [((1, 801, 10), 25), ((1, 301, 100), 25), ((5, 161, 20), 5),
                                        ((1, 500, 4), 40), ((20, 140, 64), 1)])
def test_run_based_pedestal_against_the_slot_ring(emu, tmp_path, monkeypatch, bounds, cut):
    """The two formulations of the accumulated pedestal (spectra.c:66-78) -- explicit node state
    walked line by line (slot ring; any row order) and the closed form per run of equal window
    cell (nu-sorted rows) -- on lines that sit on integer wavenumbers with shifts of both signs
    (window cells out of order between neighbouring rows), clamped windows at both ends of the
    grid, and against the oracle."""
    lines = synth.make_line_list("O2", 300, 0.5, 170.0, seed=9)
    lines["nu"] = np.sort(np.round(lines["nu"]) + np.tile([0.0, 1e-6, -1e-6, 0.5, 0.999], 60))
    lines["delta_air"] = np.tile([-0.02, 0.02, 0.0199, -0.0003, 0.015], 60)
    path = str(tmp_path / "shift.db")
    synth.write_database(path, {"O2": lines})
    gas = OracleGas(path, "O2")
    t = np.array([250.0, 296.0]); p = np.array([101325.0, 5.0e4]); x = np.array([0.209, 0.209])
    monkeypatch.delenv("EMU_PED_SLOTS", raising=False)
    runs, _ = run(emu, gas, t, p, x, bounds, 1, cut=cut)
    monkeypatch.setenv("EMU_PED_SLOTS", "1")
    slots, _ = run(emu, gas, t, p, x, bounds, 1, cut=cut)
    for layer in range(2):
        k_ref = gas.absorption(t[layer], p[layer], x[layer], *bounds, 1, cut)
        if not np.any(k_ref):
            assert not np.any(runs[layer])
            continue
        assert scaled_error(runs[layer], k_ref, bounds[2], max(cut, 1)) <= FP64_TOL
        assert scaled_error(runs[layer], slots[layer], bounds[2], max(cut, 1)) <= 1e-12


def partition(emu, cell_cost, prefix_cost, n_bands):
    cell_cost = np.ascontiguousarray(cell_cost, dtype=np.float64)
    prefix_cost = np.ascontiguousarray(prefix_cost, dtype=np.float64)
    edges = np.zeros(n_bands + 1, dtype=np.int32)
    assert emu.emu_partition_bands(cell_cost, prefix_cost, cell_cost.size, n_bands, edges) == 0
    return edges


def band_costs(cell_cost, prefix_cost, edges):
    cum = np.concatenate([[0.], np.cumsum(cell_cost)])
    return np.array([cum[hi] - cum[lo] + prefix_cost[hi] for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo])


def test_band_partition_minimises_the_largest_band(emu):
    """The arithmetic behind `lbl_gas_band_edges` (pylbl_b200/csrc/lbl_bands.h; SURVEY.md 8(e):
    contiguous bands balanced by cost): the edges cover the grid once, no band is empty while
    there are cells, and the largest band cost -- cells plus what the pedestal recurrence over
    the rows before the band's end costs -- is the smallest any contiguous partition reaches
    (checked by exhaustive search on small grids)."""
    import itertools
    rng = np.random.default_rng(11)
    for ncell, n_bands in ((7, 3), (9, 4), (12, 3), (10, 5)):
        for trial in range(6):
            cell = rng.uniform(0.1, 3.0, ncell) * (1. + 20. * (rng.random(ncell) < 0.15))
            prefix = np.concatenate([[0.], np.cumsum(rng.uniform(0., 0.6, ncell))]) * (trial % 2)
            edges = partition(emu, cell, prefix, n_bands)
            assert edges[0] == 0 and edges[-1] == ncell and np.all(np.diff(edges) >= 1)
            best = min(band_costs(cell, prefix, (0,) + cut + (ncell,)).max()
                       for cut in itertools.combinations(range(1, ncell), n_bands - 1))
            assert band_costs(cell, prefix, edges).max() <= best * (1. + 1e-12)


def test_band_partition_edge_cases(emu):
    # more bands than cells: one cell each, the rest empty at the end
    edges = partition(emu, [1., 1., 1.], [0., 0., 0., 0.], 5)
    assert list(edges) == [0, 1, 2, 3, 3, 3]
    # one band
    assert list(partition(emu, np.ones(10), np.zeros(11), 1)) == [0, 10]
    # one cell that outweighs the rest: it gets a band of its own, the others still non-empty
    cell = np.ones(8)
    cell[3] = 1000.
    edges = partition(emu, cell, np.zeros(9), 4)
    assert edges[0] == 0 and edges[-1] == 8 and np.all(np.diff(edges) >= 1)
    assert any(lo == 3 and hi == 4 for lo, hi in zip(edges[:-1], edges[1:]))
    # no cost at all (a grid beyond the line list): still a valid cover
    edges = partition(emu, np.zeros(6), np.zeros(7), 3)
    assert edges[0] == 0 and edges[-1] == 6 and np.all(np.diff(edges) >= 0)
    # uniform cells, a prefix cost that grows along the grid: later bands are narrower
    edges = partition(emu, np.ones(400), 0.5 * np.arange(401), 4)
    widths = np.diff(edges)
    assert np.all(widths[:-1] >= widths[1:]) and widths[0] > widths[-1]
