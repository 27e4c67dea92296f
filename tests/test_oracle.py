"""Pins the oracle: restatement == unmodified reference (bit for bit) where the reference
library is available, and restatement == golden vectors (generated from the reference)
everywhere."""
import numpy as np
import pytest

from oracle import (OracleGas, ReferenceGas, have_reference, oracle_regions, oracle_voigt,
                    reference_voigt)
from pylbl_b200 import synth

import golden_util

needs_reference = pytest.mark.skipif(not have_reference(),
                                     reason="oracle/_ref/libabsorption_ref.so not present")


@pytest.mark.parametrize("name", ["fixture_3gas_npv10", "fixture_co2_band_npv200",
                                  "fixture_co2_cut5_npv4"])
def test_restatement_matches_golden_bitwise(name, tmp_path):
    path, bounds, p, t, vmr, spectra, _ = golden_util.load(name, tmp_path)
    for key, k_gold in spectra.items():
        formula, cut, ped = golden_util.parse_key(key)
        gas = OracleGas(path, formula)
        for layer in range(t.size):
            k = gas.absorption(t[layer], p[layer], vmr[formula][layer], *bounds, ped, cut)
            assert np.array_equal(k, k_gold[layer]), (key, layer)


def test_voigt_known_answers():
    z = np.load(golden_util.GOLDEN / "voigt_kat.npz")
    v = z["v"]
    seen = np.zeros(7, dtype=np.int64)
    for (nu, alpha, gamma, sw), k_gold in zip(z["params"], z["k"]):
        k = np.zeros(v.size)
        oracle_voigt(v, 0, v.size - 1, nu, alpha, gamma, sw, k)
        assert np.array_equal(k, k_gold)
        seen += oracle_regions(v, 0, v.size - 1, nu, alpha, gamma)
    # The vectors exercise the Lorentz branch and every Humlicek region (voigt.c:17-186).
    assert np.all(seen > 0), seen


@needs_reference
@pytest.mark.parametrize("n_per_v", [1, 10, 100])
def test_restatement_matches_reference_live(small_db, atmosphere, n_per_v):
    for formula in ("H2O", "CO2", "O3"):
        ref = ReferenceGas(small_db, formula)
        ora = OracleGas(small_db, formula)
        for layer in range(atmosphere.t.size):
            for ped in (0, 1):
                args = (atmosphere.t[layer], atmosphere.p[layer], atmosphere.vmr[formula][layer],
                        1, 1501, n_per_v, ped)
                assert np.array_equal(ref.absorption(*args), ora.absorption(*args))


@needs_reference
def test_restatement_matches_reference_band_and_break(dense_db, small_db, atmosphere):
    # Band grid (all lines inside [v0-26, vn+26]).
    ref, ora = ReferenceGas(dense_db, "CO2"), OracleGas(dense_db, "CO2")
    args = (atmosphere.t[1], atmosphere.p[1], atmosphere.vmr["CO2"][1], 500, 851, 50, 1)
    k = ref.absorption(*args)
    assert np.any(k) and np.array_equal(k, ora.absorption(*args))
    # Early break: first row lies more than cut_off+1 below the grid -> all zeros (quirk Q1).
    ref, ora = ReferenceGas(small_db, "CO2"), OracleGas(small_db, "CO2")
    args = (atmosphere.t[1], atmosphere.p[1], atmosphere.vmr["CO2"][1], 640, 700, 10, 0)
    k = ref.absorption(*args)
    assert not np.any(k) and np.array_equal(k, ora.absorption(*args))
    assert ora.last_active == 0


@needs_reference
def test_voigt_matches_reference_live():
    rng = np.random.default_rng(7)
    v = 2000.0 + np.arange(2001) * 0.001
    for _ in range(40):
        nu = 2001.0 + rng.uniform(-0.5, 0.5)
        alpha = 10.0 ** rng.uniform(-4, -2.5)
        gamma = 10.0 ** rng.uniform(-9, -0.5)
        k0, k1 = np.zeros(v.size), np.zeros(v.size)
        reference_voigt(v, 0, v.size - 1, nu, alpha, gamma, 1e-22, k0)
        oracle_voigt(v, 0, v.size - 1, nu, alpha, gamma, 1e-22, k1)
        assert np.array_equal(k0, k1)


def test_no_tips_and_window_bookkeeping(tmp_path, atmosphere):
    path = str(tmp_path / "x.db")
    synth.write_database(path, {"CO": synth.make_line_list("CO", 40, 1.0, 300.0)}, tips=False)
    assert not np.any(OracleGas(path, "CO").absorption(250., 5e4, 1e-7, 1, 301, 10))
    path = str(tmp_path / "y.db")
    synth.write_database(path, {"CO": synth.make_line_list("CO", 40, 1.0, 300.0)})
    gas = OracleGas(path, "CO")
    gas.absorption(250., 5e4, 1e-7, 1, 301, 10, windows=True)
    win = gas.last_windows
    assert gas.last_active == 40
    assert gas.last_evals == int(np.sum(win[:, 1] - win[:, 0] + 1))
    assert np.all(win[:, 0] % 10 == 0)   # windows are integer-wavenumber aligned (quirk Q3)
