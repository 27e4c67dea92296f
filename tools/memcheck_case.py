"""A small call through every kernel of the library, for compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_case.py

Dense list (50 lines per cm-1: the far-field kernel stages several chunks per cell and copies
rounded-up byte counts, the pedestal runs through the run-based kernels), a band call, a coarse
grid (direct kernel + point-major near-zone kernel), the device-side gas sum and the continua."""
import sys
import tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pylbl_b200 import Continuum, Gas, Mixture, synth

atm = synth.fixture_atmosphere()
with tempfile.TemporaryDirectory() as tmp:
    dense = str(Path(tmp) / "dense.db")
    synth.write_database(dense, synth.config_line_lists(3, scale=1. / 3.))
    gas = Gas(dense, "CO2", devices=[0])
    fine = (500, 851, 200)
    for ped in (False, True):
        k = gas.absorption_coefficients(atm.t, atm.p, atm.vmr["CO2"], bounds=fine, remove_pedestal=ped)
        assert np.isfinite(k).all() and k.any()
        band = gas.absorption_band(atm.t, atm.p, atm.vmr["CO2"], fine, (117, 233), remove_pedestal=ped)
        assert np.array_equal(band, k[:, 117 * 200:233 * 200])
    coarse = (500, 851, 10)
    k = gas.absorption_coefficients(atm.t, atm.p, atm.vmr["CO2"], bounds=coarse, remove_pedestal=True)
    assert np.isfinite(k).all() and k.any()
    k1 = gas.absorption_coefficient(atm.t[0], atm.p[0], atm.vmr["CO2"][0],
                                    synth.grid_from_bounds(*fine), remove_pedestal=True)
    assert np.isfinite(k1).all()
    gas.close()

    small = str(Path(tmp) / "small.db")
    synth.write_database(small, synth.config_line_lists(2, scale=0.01))
    gases = ["H2O", "CO2", "O3"]
    vmr = {g: atm.vmr[g] for g in gases + ["O2"]}
    vmr["N2"] = np.full(atm.t.size, 0.78)
    mix = Mixture(small, gases)
    cont = Continuum()
    total = mix.total_absorption(atm.t, atm.p, vmr, bounds=(1, 801, 100), continuum=cont)
    assert np.isfinite(total).all() and total.any()
    mix.close()
    cont.close()
print("memcheck case done")
