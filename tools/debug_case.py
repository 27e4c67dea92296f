"""Scratch diagnostics: one failing parity case, per layer, both pedestal formulations."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from oracle import OracleGas
from pylbl_b200 import Gas, synth
from helpers import scaled_error

path = "/tmp/dbg_small.db"
synth.write_database(path, synth.config_line_lists(1, scale=0.1))
atm = synth.fixture_atmosphere()
bounds = (1, 500, 10)
for cut in (1, 2, 5):
    for runs in ("1", "0"):
        os.environ["PYLBL_B200_PEDRUNS"] = runs
        gas, ref = Gas(path, "H2O"), OracleGas(path, "H2O")
        for ped in (False, True):
            for rep in range(2):
                k = gas.absorption_coefficients(atm.t, atm.p, atm.vmr["H2O"], bounds=bounds,
                                                remove_pedestal=ped, cut_off=cut)
                errs = []
                for layer in range(4):
                    k_ref = ref.absorption(atm.t[layer], atm.p[layer], atm.vmr["H2O"][layer], *bounds, ped, cut)
                    errs.append(scaled_error(k[layer], k_ref, 10, max(cut, 1)))
                    if errs[-1] > 1e-9:
                        bad = np.argmax(np.abs(k[layer] - k_ref))
                        print("   worst point", bad, k[layer][bad], k_ref[bad])
                print(f"cut={cut} runs={runs} ped={ped} rep={rep} errs={['%.1e' % e for e in errs]}")
        gas.close()
