"""ctypes front ends of the two CPU checkers.  TEST INFRASTRUCTURE ONLY (see __init__)."""
from __future__ import annotations

import ctypes
import os
import sqlite3
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong
from pathlib import Path

import numpy as np
from numpy.ctypeslib import ndpointer

_HERE = Path(__file__).resolve().parent
_ORACLE_SO = _HERE / "liblbl_oracle.so"
_REF_SO = _HERE / "_ref" / "libabsorption_ref.so"

_f64 = ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32 = ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Runs oracle/Makefile (compiles the restatement; and the reference when
    /root/reference is present)."""
    if force or not _ORACLE_SO.exists() or \
            _ORACLE_SO.stat().st_mtime < (_HERE / "lbl_oracle.c").stat().st_mtime or \
            (not _REF_SO.exists() and os.path.isdir("/root/reference/pyLBL/c_lib")):
        subprocess.run(["make", "-C", str(_HERE), "all"], check=True,
                       stdout=subprocess.DEVNULL)


def have_reference() -> bool:
    return _REF_SO.exists()


_oracle_lib = None
_ref_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        build()
        lib = ctypes.CDLL(str(_ORACLE_SO))
        lib.lbl_oracle_absorption.argtypes = (
            [c_double] * 3 + [c_int] * 3 + [_f64] + [c_int] + [_f64] * 7 + [_i32] + [_f64] +
            [c_int, c_int] + [_f64, _f64] + [c_int, c_int] + [_f64] +
            [POINTER(c_longlong), POINTER(c_int), ctypes.c_void_p])
        lib.lbl_oracle_absorption.restype = c_int
        lib.lbl_oracle_voigt.argtypes = [_f64, c_int, c_int] + [c_double] * 4 + [_f64]
        lib.lbl_oracle_voigt.restype = None
        lib.lbl_oracle_regions.argtypes = [_f64, c_int, c_int] + [c_double] * 3 + \
            [ndpointer(np.int64, flags="C_CONTIGUOUS")]
        lib.lbl_oracle_regions.restype = None
        lib.lbl_oracle_tips.argtypes = [_f64, _f64, c_int, c_double, c_int]
        lib.lbl_oracle_tips.restype = c_double
        lib.lbl_oracle_scale_line.argtypes = [c_double] * 13 + [_f64]
        lib.lbl_oracle_scale_line.restype = None
        _oracle_lib = lib
    return _oracle_lib


def _ref():
    global _ref_lib
    if _ref_lib is None:
        build()
        if not _REF_SO.exists():
            raise RuntimeError("oracle/_ref/libabsorption_ref.so is absent (no /root/reference "
                               "here and no prebuilt copy travelled)")
        lib = ctypes.CDLL(str(_REF_SO))
        # Same argtypes as pyLBL/c_lib/gas_optics.py:68-73.
        lib.absorption.argtypes = [c_double] * 3 + [c_int] * 3 + [_f64] + [c_char_p] * 2 + \
            [c_int] * 2
        lib.absorption.restype = c_int
        lib.voigt.argtypes = [_f64, c_int, c_int] + [c_double] * 4 + [_f64]
        lib.voigt.restype = None
        _ref_lib = lib
    return _ref_lib


def grid_ints(grid):
    """pyLBL/c_lib/gas_optics.py:61-63."""
    v0 = int(round(grid[0]))
    vn = int(round(grid[-1]) + 1)
    n_per_v = int(round(1. / (grid[1] - grid[0])))
    return v0, vn, n_per_v


# --------------------------------------------------------------------------------------
# The reference itself
# --------------------------------------------------------------------------------------
class ReferenceGas(object):
    """The reference's ``Gas`` (pyLBL/c_lib/gas_optics.py:29-92) bound to oracle/_ref."""

    def __init__(self, lines_database, formula):
        self.database = getattr(lines_database, "path", lines_database)
        self.formula = formula

    def absorption_coefficient(self, temperature, pressure, volume_mixing_ratio, grid,
                               remove_pedestal=False, cut_off=25):
        v0, vn, n_per_v = grid_ints(grid)
        return self.absorption(temperature, pressure, volume_mixing_ratio, v0, vn, n_per_v,
                               remove_pedestal, cut_off)

    def absorption(self, temperature, pressure, volume_mixing_ratio, v0, vn, n_per_v,
                   remove_pedestal=False, cut_off=25):
        k = np.zeros((vn - v0) * n_per_v)
        rc = _ref().absorption(float(pressure), float(temperature), float(volume_mixing_ratio),
                               v0, vn, n_per_v, k, bytes(self.database, encoding="utf-8"),
                               bytes(self.formula, encoding="utf-8"), int(cut_off),
                               1 if remove_pedestal else 0)
        if rc != 0:
            raise ValueError("Error inside c functions.")
        return k


def reference_voigt(v, start, end, nu, alpha, gamma, sw, k):
    _ref().voigt(v, start, end, nu, alpha, gamma, sw, k)


# --------------------------------------------------------------------------------------
# The restatement
# --------------------------------------------------------------------------------------
def read_molecule(path: str, formula: str):
    """Runs the reference's four queries (absorption.c:67-71, spectral_database.c:54-56,
    112-114,142-144) with the stdlib sqlite3 module and returns plain arrays in row order.

    Returns None for the tips arrays when the molecule has no TIPS rows.
    """
    con = sqlite3.connect(path)
    cur = con.cursor()
    row = cur.execute("select molecule from molecule_alias where alias == ?",
                      (formula,)).fetchone()
    if row is None:
        con.close()
        raise ValueError("Error inside c functions.")
    mid = int(row[0])
    tips = cur.execute("select isotopologue_id, temperature, data from tips "
                       f"where molecule_id == {mid}").fetchall()
    mass = np.zeros(32)
    for isoid, m in cur.execute(f"select isoid, mass from isotopologue where molecule_id == {mid}"):
        isoid = 10 if isoid == 0 else isoid
        mass[isoid - 1] = m
    rows = cur.execute("select nu, sw, gamma_air, gamma_self, n_air, elower, delta_air, "
                       f"local_iso_id from transition where molecule_id == {mid}").fetchall()
    con.close()
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    out = dict(nu=a[:, 0].copy(), sw=a[:, 1].copy(), gamma_air=a[:, 2].copy(),
               gamma_self=a[:, 3].copy(), n_air=a[:, 4].copy(), elower=a[:, 5].copy(),
               delta_air=a[:, 6].copy(), local_iso_id=a[:, 7].astype(np.int32), mass=mass)
    if tips:
        t = np.asarray(tips, dtype=np.float64)
        num_iso = 1 + int(np.count_nonzero(np.diff(t[:, 0])))
        out.update(num_iso=num_iso, num_t=t.shape[0] // num_iso,
                   tips_t=t[:, 1].copy(), tips_q=t[:, 2].copy())
    else:
        out.update(num_iso=0, num_t=0, tips_t=None, tips_q=None)
    return out


class OracleGas(object):
    """``Gas``-shaped front end of oracle/lbl_oracle.c."""

    def __init__(self, lines_database, formula):
        self.database = getattr(lines_database, "path", lines_database)
        self.formula = formula
        self.data = read_molecule(self.database, formula)
        self.last_evals = 0
        self.last_active = 0
        self.last_windows = None

    def absorption_coefficient(self, temperature, pressure, volume_mixing_ratio, grid,
                               remove_pedestal=False, cut_off=25):
        v0, vn, n_per_v = grid_ints(grid)
        return self.absorption(temperature, pressure, volume_mixing_ratio, v0, vn, n_per_v,
                               remove_pedestal, cut_off)

    def absorption(self, temperature, pressure, volume_mixing_ratio, v0, vn, n_per_v,
                   remove_pedestal=False, cut_off=25, windows=False):
        d = self.data
        n = (vn - v0) * n_per_v
        k = np.zeros(n)
        if d["tips_t"] is None:
            return k  # absorption.c:53-59
        work = np.empty(n)
        evals = c_longlong(0)
        active = c_int(0)
        win = None
        win_ptr = None
        if windows:
            win = np.full(2 * d["nu"].size, -1, dtype=np.int32)
            win_ptr = win.ctypes.data_as(ctypes.c_void_p)
        _oracle().lbl_oracle_absorption(
            float(pressure), float(temperature), float(volume_mixing_ratio), v0, vn, n_per_v, k,
            d["nu"].size, d["nu"], d["sw"], d["gamma_air"], d["gamma_self"], d["n_air"],
            d["elower"], d["delta_air"], d["local_iso_id"], d["mass"], d["num_iso"], d["num_t"],
            d["tips_t"], d["tips_q"], int(cut_off), 1 if remove_pedestal else 0, work,
            ctypes.byref(evals), ctypes.byref(active), win_ptr)
        self.last_evals = int(evals.value)
        self.last_active = int(active.value)
        self.last_windows = None if win is None else win.reshape(-1, 2)
        return k


def oracle_voigt(v, start, end, nu, alpha, gamma, sw, k):
    _oracle().lbl_oracle_voigt(v, start, end, nu, alpha, gamma, sw, k)


def oracle_regions(v, start, end, nu, alpha, gamma):
    hist = np.zeros(7, dtype=np.int64)
    _oracle().lbl_oracle_regions(v, start, end, nu, alpha, gamma, hist)
    return hist


def oracle_tips(t, q, num_t, temperature, iso):
    return _oracle().lbl_oracle_tips(t, q, num_t, float(temperature), int(iso))


def oracle_scale_line(temperature, pressure, abundance, nu, sw, gamma_air, gamma_self, n_air,
                      elower, delta_air, mass, q_ref, q_t):
    out = np.zeros(4)
    _oracle().lbl_oracle_scale_line(float(temperature), float(pressure), float(abundance),
                                    float(nu), float(sw), float(gamma_air), float(gamma_self),
                                    float(n_air), float(elower), float(delta_air), float(mass),
                                    float(q_ref), float(q_t), out)
    return out
