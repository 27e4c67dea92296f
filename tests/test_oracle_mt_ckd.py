"""The numpy restatement of the MT-CKD continuum (oracle/mt_ckd.py) against vectors produced by
the reference's own modules (tests/golden/make_mt_ckd_golden.py), including the reference's own
known-answer test (tests/test_mt_ckd.py:15-46 there)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import mt_ckd

GOLDEN = np.load(Path(__file__).resolve().parent / "golden" / "mt_ckd_reference.npz")
NAMES = ["CO2", "H2OForeign", "H2OSelf", "N2", "O2", "O3"]
GASES = ["H2O", "CO2", "O3", "N2O", "CH4", "CO", "O2", "N2"]

# tests/test_mt_ckd.py:15-27 of the reference
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


def vmr_of(layer):
    return {g: float(GOLDEN["vmr_" + g][layer]) for g in GASES}


@pytest.mark.parametrize("name", NAMES)
def test_band_sums_match_the_references_known_answers(name):
    """As the reference's test calls it: band.spectra(T, p, vmr) with p = 98388 (Pa passed where
    the band expects mb -- that is how the recorded numbers were made)."""
    layer = 3
    t, p = float(GOLDEN["temperature"][layer]), float(GOLDEN["pressure"][layer])
    bands = mt_ckd.OracleContinuum(name).bands
    assert len(bands) == len(KNOWN_ANSWERS[name])
    for (w, band), want, stored in zip(bands, KNOWN_ANSWERS[name], GOLDEN["band_sums_" + name]):
        got = float(np.sum(band(t, p, vmr_of(layer))))
        assert got == pytest.approx(want)                 # the reference's own tolerance (1e-6)
        assert got == pytest.approx(stored, rel=1e-13)    # what its code returned here


@pytest.mark.parametrize("grid_name", ["coarse", "bandhead", "uv"])
@pytest.mark.parametrize("name", NAMES)
def test_interpolated_continua_match_the_reference(name, grid_name):
    grid = GOLDEN["grid_" + grid_name]
    want = GOLDEN[f"{name}_{grid_name}"]
    oracle = mt_ckd.OracleContinuum(name)
    for row, layer in enumerate(GOLDEN["layers"]):
        got = oracle.spectra(float(GOLDEN["temperature"][layer]), float(GOLDEN["pressure"][layer]),
                             vmr_of(int(layer)), grid)
        scale = np.abs(want[row]).max()
        if scale == 0.:
            assert not got.any()
        else:
            assert np.abs(got - want[row]).max() <= 1e-13 * scale
    # the three grids together see every continuum
    assert any(GOLDEN[f"{name}_{g}"].any() for g in ("coarse", "bandhead", "uv"))


def test_driver_attaches_the_references_continua():
    assert mt_ckd.continua_of("H2O") == ["H2OForeign", "H2OSelf"]     # spectroscopy.py:58-61
    assert mt_ckd.continua_of("CO2") == ["CO2"] and mt_ckd.continua_of("CH4") == []
