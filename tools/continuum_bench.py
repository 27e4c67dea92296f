"""HBM roofline of the continuum pass (K5b) at BASELINE configs[4] shape: 256 columns x 60 layers,
grid 1-5000 cm-1 @0.1 (50 000 points), every continuum of the 7-gas atmosphere (+ N2), added
into the gas-sum accumulator on the device.  python tools/continuum_bench.py [columns]"""
import ctypes
import json
import sys
from ctypes import c_void_p
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pylbl_b200 import Continuum, _lib, continua_of, synth

columns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cols = [synth.standard_column(60, column=c) for c in range(columns)]
t = np.concatenate([c.t for c in cols])
p = np.concatenate([c.p for c in cols])
gases = ["H2O", "CO2", "O3", "N2O", "CO", "CH4", "O2"]
vmr = {g: np.concatenate([c.vmr[g] for c in cols]) for g in gases}
vmr["N2"] = np.full(t.size, 0.78)
bounds = synth.config_grid(5)
n = (bounds[1] - bounds[0]) * bounds[2]
lib = _lib.library()
mix = c_void_p()
lib.lbl_mix_open(0, int(t.size), n, ctypes.byref(mix))
cont = Continuum(0)
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6544.0
rows = []
for rep in range(3):
    for formula in vmr:
        for name in continua_of(formula):
            cont.spectra(name, t, p, vmr, bounds=bounds, mix=mix)
            bands_ms, apply_ms = cont.last_ms()
            if rep == 2:
                nbytes = 16.0 * t.size * n          # accumulator read + written, 8 B each per point
                rows.append({"continuum": name, "layers": int(t.size), "points": n, "bands_ms": bands_ms,
                             "apply_ms": apply_ms, "algorithmic_bytes": nbytes,
                             "achieved_gbs": nbytes / (apply_ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": nbytes / (apply_ms * 1e-3) / 1e9 / peak})
names = [name for formula in vmr for name in continua_of(formula)]
for rep in range(3):
    cont.spectra(names, t, p, vmr, bounds=bounds, mix=mix)
bands_ms, apply_ms = cont.last_ms()
nbytes = 16.0 * t.size * n
rows.append({"continuum": "+".join(names) + " (one pass)", "layers": int(t.size), "points": n,
             "bands_ms": bands_ms, "apply_ms": apply_ms, "algorithmic_bytes": nbytes,
             "achieved_gbs": nbytes / (apply_ms * 1e-3) / 1e9,
             "frac_of_hbm_peak": nbytes / (apply_ms * 1e-3) / 1e9 / peak})
for r in rows:
    print(json.dumps(r))
print(json.dumps({"kernel": "lbl::continuum_apply_kernel<true>", "bound": "hbm", "peak_gbs": peak,
                  "separate_passes_apply_ms": sum(r["apply_ms"] for r in rows[:-1]),
                  "one_pass_apply_ms": rows[-1]["apply_ms"], "one_pass_frac": rows[-1]["frac_of_hbm_peak"],
                  "mean_frac_separate": float(np.mean([r["frac_of_hbm_peak"] for r in rows[:-1]]))}))
