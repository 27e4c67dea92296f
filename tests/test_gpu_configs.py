"""BASELINE.json configurations at full size on the GPU.

Where the oracle finishes in seconds it is the judge (configs 1, 3, the small gases of 2, one
band of 4).  Where it does not, size-independent properties are: superposition (the spectrum
of a line list is the sum of the spectra of a partition of it), band decomposition (a band
computed from a database pre-filtered to [v0-26, vn+26] equals the same slice of the wide
computation), and the evaluation count (sum of window lengths), which is known in closed
form from the window arithmetic.
"""
import numpy as np
import pytest

from oracle import OracleGas
from pylbl_b200 import Gas, synth

from helpers import FP64_TOL, relative_error, scaled_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def db_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("configs")


def test_config1_full_size(db_dir):
    """configs[0]: single layer (98388 Pa, 289 K), H2O+CO2+O3, ~50k lines, 1-5000 @0.1."""
    path = str(db_dir / "c1.db")
    synth.write_database(path, synth.config_line_lists(1))
    atm = synth.fixture_atmosphere()
    bounds = synth.config_grid(1)
    layer = 3
    for formula in ("H2O", "CO2", "O3"):
        gas, ref = Gas(path, formula), OracleGas(path, formula)
        for ped in (False, True):
            args = (atm.t[layer], atm.p[layer], atm.vmr[formula][layer])
            k = gas.absorption_coefficients([args[0]], [args[1]], [args[2]], bounds=bounds,
                                            remove_pedestal=ped)[0]
            k_ref = ref.absorption(*args, *bounds, ped)
            assert scaled_error(k, k_ref, bounds[2]) <= FP64_TOL
            if not ped:
                assert relative_error(k, k_ref) <= FP64_TOL
            assert gas.last_stats[0]["evals"] == ref.last_evals


def test_config2_small_gases_and_superposition(db_dir):
    """configs[1]: 60-layer column, 7 gases, 1-5000 @0.01 (500 000 points per spectrum)."""
    lists = synth.config_line_lists(2)
    path = str(db_dir / "c2.db")
    synth.write_database(path, lists)
    col = synth.standard_column(60)
    bounds = synth.config_grid(2)
    # direct parity for the two short line lists, a few layers, pedestal on
    for formula in ("CO", "O2"):
        gas, ref = Gas(path, formula), OracleGas(path, formula)
        k = gas.absorption_coefficients(col.t, col.p, col.vmr[formula], bounds=bounds,
                                        remove_pedestal=True)
        for layer in (0, 31, 59):
            k_ref = ref.absorption(col.t[layer], col.p[layer], col.vmr[formula][layer], *bounds, True)
            assert scaled_error(k[layer], k_ref, bounds[2]) <= FP64_TOL
    # superposition at full size: H2O = odd lines + even lines (no pedestal: it is not linear)
    h2o = lists["H2O"]
    halves = []
    for parity in (0, 1):
        sub = {key: val[parity::2] for key, val in h2o.items()}
        p = str(db_dir / f"c2_h2o_{parity}.db")
        synth.write_database(p, {"H2O": sub})
        halves.append(Gas(p, "H2O"))
    layers = [0, 40]
    t, pr, x = col.t[layers], col.p[layers], col.vmr["H2O"][layers]
    whole = Gas(path, "H2O").absorption_coefficients(t, pr, x, bounds=bounds)
    parts = sum(g.absorption_coefficients(t, pr, x, bounds=bounds) for g in halves)
    assert np.all(whole > 0)
    # not bit-equal: lines are paired differently, and each reciprocal carries its own
    # ~1e-12 Newton residual
    assert relative_error(parts, whole) <= 1e-10


def test_config3_dense_band(db_dir):
    """configs[2]: CO2 15-um band 500-850 cm-1 @0.0005 (702 000 points), ~60k lines."""
    path = str(db_dir / "c3.db")
    synth.write_database(path, synth.config_line_lists(3))
    col = synth.standard_column(60)
    bounds = synth.config_grid(3)
    gas, ref = Gas(path, "CO2"), OracleGas(path, "CO2")
    layer = 25
    k = gas.absorption_coefficients(col.t[layer:layer + 1], col.p[layer:layer + 1],
                                    col.vmr["CO2"][layer:layer + 1], bounds=bounds,
                                    remove_pedestal=True)[0]
    k_ref = ref.absorption(col.t[layer], col.p[layer], col.vmr["CO2"][layer], *bounds, True)
    assert scaled_error(k, k_ref, bounds[2]) <= FP64_TOL
    assert gas.last_stats[0]["evals"] == ref.last_evals


def test_config4_million_lines_bands(db_dir):
    """configs[3]: ~1M lines, 10-3500 cm-1 @0.001 (3 491 000 points): band decomposition."""
    lines = synth.config_line_lists(4)["XX"]
    path = str(db_dir / "c4.db")
    synth.write_database(path, {"XX": lines})
    col = synth.standard_column(60)
    layer = 12
    t, p, x = col.t[layer:layer + 1], col.p[layer:layer + 1], col.vmr["XX"][layer:layer + 1]
    v0, vn, npv = synth.config_grid(4)
    wide = Gas(path, "XX")
    k_wide = wide.absorption_coefficients(t, p, x, bounds=(v0, vn, npv))[0]
    stats = wide.last_stats[0]
    assert stats["n_active"] == lines["nu"].size
    # closed-form evaluation count from the window arithmetic (spectra.c:48-62)
    n = (vn - v0) * npv
    cb = np.floor(lines["nu"] + p[0] * 9.86923e-6 * lines["delta_air"]) - v0
    s = np.maximum((cb - 25) * npv, 0)
    e = np.minimum((cb + 26) * npv, n - 1)
    assert stats["evals"] == int(np.sum((e - s + 1)[(cb - 25) * npv < n]))
    for band in ((1000, 1012), (2290, 2300)):
        keep = (lines["nu"] >= band[0] - 26) & (lines["nu"] <= band[1] + 26)
        sub = {key: val[keep] for key, val in lines.items()}
        bp = str(db_dir / f"c4_{band[0]}.db")
        synth.write_database(bp, {"XX": sub})
        bounds = (band[0], band[1], npv)
        k_band = Gas(bp, "XX").absorption_coefficients(t, p, x, bounds=bounds)[0]
        sl = slice((band[0] - v0) * npv, (band[1] - v0) * npv)
        assert relative_error(k_band, k_wide[sl]) <= 1e-10
        k_ref = OracleGas(bp, "XX").absorption(t[0], p[0], x[0], *bounds)
        assert relative_error(k_band, k_ref) <= FP64_TOL
