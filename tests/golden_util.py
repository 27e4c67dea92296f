"""Loads tests/golden/*.npz and rebuilds their sqlite databases."""
from pathlib import Path

import numpy as np

from pylbl_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"
LINE_KEYS = ("nu", "sw", "gamma_air", "gamma_self", "n_air", "elower", "delta_air", "local_iso_id")


def load(name, tmp_dir):
    """Returns (db_path, bounds, atmosphere arrays, {key: spectra}, formulas)."""
    z = np.load(GOLDEN / f"{name}.npz")
    formulas = sorted({k.split("_")[1] for k in z.files if k.startswith("lines_")})
    line_lists = {f: {key: z[f"lines_{f}_{key}"] for key in LINE_KEYS} for f in formulas}
    path = str(Path(tmp_dir) / f"{name}.db")
    synth.write_database(path, line_lists)
    spectra = {k: z[k] for k in z.files if k.startswith("k_")}
    vmr = {f: z[f"vmr_{f}"] for f in formulas}
    return path, tuple(int(x) for x in z["bounds"]), z["p"], z["t"], vmr, spectra, formulas


def parse_key(key):
    """'k_CO2_cut25_ped1' -> ('CO2', 25, 1)."""
    _, formula, cut, ped = key.split("_")
    return formula, int(cut[3:]), int(ped[3:])
