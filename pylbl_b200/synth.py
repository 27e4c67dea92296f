"""Synthetic HITRAN-shaped inputs: sqlite line databases, atmospheres and grids.

There is no network in the build or benchmark environment, so the HITRAN download that
the reference performs (pyLBL/database.py:148-210, pyLBL/webapi/hitran_api.py:121-185)
cannot run.  This module writes sqlite files with the *same tables and columns* the
reference C reader queries (pyLBL/c_lib/absorption.c:67-71,
pyLBL/c_lib/spectral_database.c:54-56,112-114,142-144; schema pyLBL/database.py:427-469)
so that the reference library and the CUDA backend consume byte-identical inputs.

Ordering invariants honoured:
  * ``transition`` rows are inserted in ascending ``nu`` (HITRAN order);
  * ``tips`` rows are inserted isotopologue-major / temperature-minor, on an integer
    1-K temperature axis, with 0-based ``isotopologue_id`` (pyLBL/database.py:117-127);
  * every ``local_iso_id`` is in ``1..num_iso``.

Everything is seeded; the same call always produces the same file contents.
"""
from __future__ import annotations

import os
import sqlite3
from collections import namedtuple

import numpy as np

# HITRAN molecule numbers and the masses [g mol-1] of the first four isotopologues.
MOLECULES = {
    "H2O": dict(id=1, mass=[18.010565, 20.014811, 19.01478, 19.01674],
                bands=[(100., 180.), (1595., 160.), (3750., 170.), (5300., 150.)]),
    "CO2": dict(id=2, mass=[43.98983, 44.993185, 45.994076, 44.994045],
                bands=[(667., 45.), (960., 30.), (1064., 30.), (2349., 50.), (3660., 60.)]),
    "O3": dict(id=3, mass=[47.984745, 49.988991, 49.988991, 48.98896],
               bands=[(60., 40.), (701., 35.), (1042., 40.), (1103., 30.), (2110., 40.)]),
    "N2O": dict(id=4, mass=[44.001062, 44.998096, 44.998096, 46.005308],
                bands=[(589., 30.), (1285., 40.), (2224., 40.), (2563., 40.), (3480., 40.)]),
    "CO": dict(id=5, mass=[27.994915, 28.99827, 29.999161, 28.99913],
               bands=[(40., 30.), (2143., 60.), (4260., 60.)]),
    "CH4": dict(id=6, mass=[16.0313, 17.034655, 17.037475, 18.04083],
                bands=[(1306., 60.), (1534., 50.), (3019., 80.), (4340., 100.)]),
    "O2": dict(id=7, mass=[31.98983, 33.994076, 32.994045, 31.98983],
               bands=[(60., 50.), (1556., 50.)]),
    # A generic heavy absorber used for the ~1M-line stress list (BASELINE config 4).
    "XX": dict(id=60, mass=[44.0, 45.0, 46.0, 47.0],
               bands=[(400., 200.), (1200., 300.), (2300., 250.), (3100., 200.)]),
}

ISO_PROBABILITY = np.array([0.9, 0.06, 0.03, 0.01])
TIPS_T_MIN, TIPS_T_MAX = 1, 1000  # integer kelvin axis, 1-K spacing

Atmosphere = namedtuple("Atmosphere", ["p", "t", "vmr"])
"""Plain arrays: p [Pa] (n_layers,), t [K] (n_layers,), vmr {formula: (n_layers,)}."""


def _seed_for(formula: str, seed: int) -> int:
    return int(seed) * 1009 + sum(ord(c) * (i + 1) for i, c in enumerate(formula))


def make_line_list(formula: str, n_lines: int, nu_min: float, nu_max: float, seed: int = 0):
    """Draws a HITRAN-shaped line list, sorted by line centre.

    Distributions follow SURVEY.md section 8(d): band clumps plus a uniform floor for
    ``nu``; log-uniform strengths; air/self widths, exponent, shift and lower-state
    energy in HITRAN-typical ranges.

    Returns:
        dict of float64 arrays ``nu, sw, gamma_air, gamma_self, n_air, elower, delta_air``
        and int32 ``local_iso_id``, all of length ``n_lines``.
    """
    rng = np.random.default_rng(_seed_for(formula, seed))
    bands = MOLECULES[formula]["bands"]
    n_clump = int(0.7 * n_lines)
    centre = rng.integers(0, len(bands), size=n_clump)
    mu = np.array([b[0] for b in bands])[centre]
    sigma = np.array([b[1] for b in bands])[centre]
    nu_clump = mu + sigma * rng.standard_normal(n_clump)
    bad = (nu_clump < nu_min) | (nu_clump > nu_max)
    nu_clump[bad] = rng.uniform(nu_min, nu_max, size=int(bad.sum()))
    nu_floor = rng.uniform(nu_min, nu_max, size=n_lines - n_clump)
    nu = np.sort(np.concatenate([nu_clump, nu_floor]))
    # HITRAN stores six decimals.
    nu = np.round(nu, 6)
    nu.sort()
    lines = dict(
        nu=nu,
        sw=10.0 ** rng.uniform(-30.0, -19.0, size=n_lines),
        gamma_air=np.round(rng.uniform(0.03, 0.11, size=n_lines), 4),
        gamma_self=np.round(rng.uniform(0.05, 0.5, size=n_lines), 3),
        n_air=np.round(rng.uniform(0.4, 0.8, size=n_lines), 2),
        elower=np.round(rng.uniform(0.0, 6000.0, size=n_lines), 4),
        delta_air=np.round(np.clip(0.003 * rng.standard_normal(n_lines), -0.02, 0.02), 6),
        local_iso_id=rng.choice(np.arange(1, 5), size=n_lines,
                                p=ISO_PROBABILITY).astype(np.int32),
    )
    return lines


def tips_table(formula: str):
    """Smooth synthetic partition sums Q_iso(T) on the integer axis 1..1000 K.

    Returns:
        temperature (num_t,), data (num_iso, num_t).
    """
    t = np.arange(TIPS_T_MIN, TIPS_T_MAX + 1, dtype=np.float64)
    mol = MOLECULES[formula]
    rows = []
    for iso in range(len(mol["mass"])):
        a = 0.03 * (1.0 + 0.35 * iso) * (1.0 + 0.01 * mol["id"])
        rows.append(a * t ** 1.5 * (1.0 + 2.0e-4 * (iso + 1) * t) + 1.0)
    return t, np.asarray(rows)


def write_database(path: str, line_lists: dict, tips: bool = True) -> str:
    """Writes a sqlite file the reference C reader and the CUDA packer can both open.

    Args:
        path: output file (overwritten).
        line_lists: {formula: line-list dict from make_line_list}.
        tips: write the ``tips`` table (False reproduces the "no TIPS data" path,
              pyLBL/c_lib/absorption.c:53-59).
    """
    if os.path.exists(path):
        os.remove(path)
    con = sqlite3.connect(path)
    cur = con.cursor()
    cur.executescript(
        """
        create table molecule (id integer primary key, stoichiometric_formula text,
                               ordinary_formula text, common_name text);
        create table isotopologue (id integer primary key, molecule_id integer,
                                   isoid integer, iso_name text, abundance float,
                                   mass float);
        create table molecule_alias (id integer primary key autoincrement, alias text,
                                     molecule integer);
        create table transition (id integer primary key autoincrement,
                                 global_iso_id integer, molecule_id integer,
                                 local_iso_id integer, nu float, sw float,
                                 gamma_air float, gamma_self float, n_air float,
                                 delta_air float, elower float);
        create table tips (id integer primary key autoincrement, molecule_id integer,
                           isotopologue_id integer, temperature float, data float);
        """
    )
    for formula, lines in line_lists.items():
        mol = MOLECULES[formula]
        mid = mol["id"]
        cur.execute("insert into molecule values (?, ?, ?, ?)", (mid, formula, formula, formula))
        cur.execute("insert into molecule_alias (alias, molecule) values (?, ?)", (formula, mid))
        for isoid, mass in enumerate(mol["mass"], start=1):
            cur.execute("insert into isotopologue (molecule_id, isoid, iso_name, abundance, mass)"
                        " values (?, ?, ?, ?, ?)",
                        (mid, isoid, f"{formula}-{isoid}", float(ISO_PROBABILITY[isoid - 1]),
                         float(mass)))
        n = len(lines["nu"])
        rows = zip([mid] * n, lines["local_iso_id"].tolist(), lines["nu"].tolist(),
                   lines["sw"].tolist(), lines["gamma_air"].tolist(),
                   lines["gamma_self"].tolist(), lines["n_air"].tolist(),
                   lines["delta_air"].tolist(), lines["elower"].tolist())
        cur.executemany(
            "insert into transition (molecule_id, local_iso_id, nu, sw, gamma_air, gamma_self,"
            " n_air, delta_air, elower) values (?, ?, ?, ?, ?, ?, ?, ?, ?)", rows)
        if tips:
            t, q = tips_table(formula)
            for iso in range(q.shape[0]):
                cur.executemany(
                    "insert into tips (molecule_id, isotopologue_id, temperature, data)"
                    " values (?, ?, ?, ?)",
                    zip([mid] * t.size, [iso] * t.size, t.tolist(), q[iso].tolist()))
    con.commit()
    con.close()
    return path


# --------------------------------------------------------------------------------------
# Atmospheres
# --------------------------------------------------------------------------------------
def fixture_atmosphere() -> Atmosphere:
    """The reference's own 4-layer test atmosphere (tests/conftest.py:54-78)."""
    p = np.asarray([117., 1032., 11419., 98388.])
    t = np.asarray([269.01, 227.74, 203.37, 288.99])
    vmr = {
        "H2O": np.asarray([5.244536e-06, 4.763972e-06, 3.039952e-06, 6.637074e-03]),
        "CO2": np.asarray([0.00036, 0.00036, 0.00036, 0.00035999]),
        "O3": np.asarray([2.936688e-06, 7.415223e-06, 2.609510e-07, 6.859128e-08]),
        "N2O": np.asarray([1.050928e-08, 1.319584e-07, 2.895416e-07, 3.199949e-07]),
        "CH4": np.asarray([2.947482e-07, 8.817705e-07, 1.588336e-06, 1.700002e-06]),
        "CO": np.asarray([3.621464e-08, 1.761450e-08, 3.315927e-08, 1.482969e-07]),
        "O2": np.asarray([0.209, 0.209, 0.2090003, 0.208996]),
        "XX": np.asarray([4.0e-4, 4.0e-4, 4.0e-4, 4.0e-4]),
    }
    return Atmosphere(p=p, t=t, vmr=vmr)


def standard_column(n_layers: int = 60, column: int = 0, seed: int = 0) -> Atmosphere:
    """A clear-sky column: p log-spaced 101325 -> 10 Pa, piecewise-linear T(z).

    ``column`` > 0 applies a seeded perturbation (T +-15 K, H2O x/ 3), keeping
    150 K < T < 330 K (SURVEY.md section 8(d) "Atmospheres").
    """
    p = np.exp(np.linspace(np.log(101325.0), np.log(10.0), n_layers))
    z = -7.0 * np.log(p / 101325.0)  # km, 7 km scale height
    t = np.where(z < 11.0, 288.15 - 6.5 * z,
                 np.where(z < 20.0, 216.65,
                          np.where(z < 32.0, 216.65 + 1.0 * (z - 20.0),
                                   np.where(z < 47.0, 228.65 + 2.8 * (z - 32.0),
                                            np.maximum(270.65 - 2.0 * (z - 47.0), 180.0)))))
    h2o = np.maximum(6.6e-3 * np.exp(-z / 2.0), 5.0e-6)
    o3 = 6.0e-8 + 8.0e-6 * np.exp(-0.5 * ((np.log(p) - np.log(3000.0)) / 0.9) ** 2)
    if column > 0:
        rng = np.random.default_rng(7919 * int(seed) + int(column))
        t = t + rng.uniform(-15.0, 15.0) + 2.0 * rng.standard_normal(n_layers)
        h2o = h2o * 3.0 ** rng.uniform(-1.0, 1.0)
    t = np.clip(np.round(t, 2), 151.0, 329.0)
    vmr = {
        "H2O": h2o,
        "CO2": np.full(n_layers, 3.6e-4),
        "O3": o3,
        "N2O": np.full(n_layers, 3.2e-7) * np.minimum(1.0, (p / 2.0e4) ** 0.3),
        "CO": np.full(n_layers, 1.5e-7) * np.minimum(1.0, (p / 5.0e4) ** 0.2),
        "CH4": np.full(n_layers, 1.7e-6) * np.minimum(1.0, (p / 1.0e4) ** 0.25),
        "O2": np.full(n_layers, 0.209),
        "XX": np.full(n_layers, 4.0e-4),
    }
    return Atmosphere(p=p, t=t, vmr=vmr)


# --------------------------------------------------------------------------------------
# BASELINE.json configurations made concrete
# --------------------------------------------------------------------------------------
CONFIG2_SHARES = {"H2O": 70000, "CO2": 90000, "O3": 80000, "N2O": 30000, "CO": 2000,
                  "CH4": 75000, "O2": 3000}


def grid_from_bounds(v0: int, vn: int, n_per_v: int) -> np.ndarray:
    """The user-side grid ``arange(v0, vn-1+..., 1/n_per_v)`` whose ctypes ints are
    (v0, vn, n_per_v) under pyLBL/c_lib/gas_optics.py:61-63."""
    n = (vn - 1 - v0) * n_per_v
    return v0 + np.arange(n) / float(n_per_v)


def config_line_lists(config: int, scale: float = 1.0, seed: int = 0) -> dict:
    """Line lists of BASELINE.json config 1..4 (``scale`` shrinks line counts for tests)."""
    if config == 1:
        shares = {"H2O": 20000, "CO2": 20000, "O3": 10000}
        lo, hi = 0.5, 5025.0
    elif config in (2, 5):
        shares = CONFIG2_SHARES
        lo, hi = 0.5, 5025.0
    elif config == 3:
        shares = {"CO2": 60000}
        lo, hi = 474.5, 875.5  # inside [v0-26, vn+26] for v0=500, vn=851 (quirk Q1)
    elif config == 4:
        shares = {"XX": 1000000}
        lo, hi = 0.5, 3526.0  # inside [10-26, 3501+26]
    else:
        raise ValueError(f"unknown config {config}")
    out = {}
    for formula, n in shares.items():
        n = max(int(round(n * scale)), 8)
        out[formula] = make_line_list(formula, n, lo, hi, seed=1000 + config + seed)
    return out


def config_grid(config: int):
    """(v0, vn, n_per_v) of BASELINE.json config 1..5."""
    return {1: (1, 5001, 10), 2: (1, 5001, 100), 3: (500, 851, 2000),
            4: (10, 3501, 1000), 5: (1, 5001, 10)}[config]
