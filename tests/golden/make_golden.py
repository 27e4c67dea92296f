"""Generates tests/golden/*.npz by running the UNMODIFIED reference C library
(oracle/_ref/libabsorption_ref.so, compiled from /root/reference by oracle/Makefile) on
synthetic HITRAN-shaped databases.  Run in the build container (where /root/reference
exists):

    python tests/golden/make_golden.py

Each fixture stores the line list (so the database can be rebuilt bit-identically on a
machine without the reference), the call arguments, and the spectra the reference
produced.  The reference's own tests pin this path only against a downloaded HITRAN
database (tests/test_gas_optics.py:17-19), which cannot be fetched offline; these fixtures
are the offline known-answer vectors.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ReferenceGas, have_reference, reference_voigt  # noqa: E402
from pylbl_b200 import synth  # noqa: E402

HERE = Path(__file__).resolve().parent


def spectra_fixture(name, line_lists, bounds, cut_offs=(25,), tmp="/tmp/golden.db"):
    synth.write_database(tmp, line_lists)
    atm = synth.fixture_atmosphere()
    v0, vn, n_per_v = bounds
    out = {"bounds": np.asarray(bounds), "p": atm.p, "t": atm.t}
    for formula, lines in line_lists.items():
        for key, value in lines.items():
            out[f"lines_{formula}_{key}"] = value
        out[f"vmr_{formula}"] = atm.vmr[formula]
        ref = ReferenceGas(tmp, formula)
        for cut in cut_offs:
            for ped in (0, 1):
                k = np.stack([ref.absorption(atm.t[i], atm.p[i], atm.vmr[formula][i], v0, vn,
                                             n_per_v, ped, cut) for i in range(atm.t.size)])
                out[f"k_{formula}_cut{cut}_ped{ped}"] = k
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith("k_")})


def voigt_fixture():
    """Known-answer vectors for voigt() alone (pyLBL/c_lib/voigt.c:4-191): every region."""
    rng = np.random.default_rng(42)
    v = 1000.0 + np.arange(4001) * 0.0005     # +-1 cm-1 around the centre at 5e-4 spacing
    cases = []
    for alpha, gamma in [(8e-4, 7e-2), (8e-4, 5e-3), (8e-4, 4e-4), (8e-4, 3e-5), (8e-4, 5e-10),
                         (2e-3, 2e-3), (1e-4, 1.2e-3), (3e-3, 1e-5), (5e-4, 3.6e-2),
                         (1e-3, 8.6e-2)]:
        nu = 1001.0 + rng.uniform(-0.3, 0.3)
        sw = 10.0 ** rng.uniform(-26, -20)
        k = np.zeros(v.size)
        reference_voigt(v, 0, v.size - 1, nu, alpha, gamma, sw, k)
        cases.append((nu, alpha, gamma, sw, k))
    np.savez_compressed(HERE / "voigt_kat.npz", v=v,
                        params=np.asarray([c[:4] for c in cases]),
                        k=np.stack([c[4] for c in cases]))
    print("voigt_kat", len(cases))


if __name__ == "__main__":
    if not have_reference():
        raise SystemExit("oracle/_ref/libabsorption_ref.so is absent; run `make -C oracle`")
    voigt_fixture()
    spectra_fixture("fixture_3gas_npv10",
                    {f: synth.make_line_list(f, n, 0.5, 425.0, seed=21)
                     for f, n in (("H2O", 300), ("CO2", 300), ("O3", 200))},
                    (1, 401, 10))
    spectra_fixture("fixture_co2_cut5_npv4",
                    {"CO2": synth.make_line_list("CO2", 300, 0.5, 425.0, seed=23)},
                    (1, 401, 4), cut_offs=(5,))
    spectra_fixture("fixture_co2_band_npv200",
                    {"CO2": synth.make_line_list("CO2", 500, 614.5, 725.5, seed=22)},
                    (640, 701, 200))
