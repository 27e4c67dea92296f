"""Gas-summed line absorption on the device.

pyLBL's driver turns each gas's cross-sections into absorption coefficients on the host,
``beta = n * k[:grid.size]`` with ``n = p*x/(kB*T)`` (pyLBL/spectroscopy.py:18-29,181-191),
and sums the gases when ``output_format="total"`` (:225-234).  With every spectrum coming
back over PCIe that sum costs seven device-to-host copies per column.  ``Mixture`` keeps the
per-gas spectra on the GPU, applies the number densities there and returns one array
(``lbl_gas_submit_mix``: nothing waits on the host between the gases).
"""
from __future__ import annotations

import ctypes
import os
import time
from ctypes import c_void_p

import numpy as np

from . import _lib
from .gas_optics import Gas, default_device, grid_to_ints

KB = 1.38064852e-23  # Boltzmann constant [J K-1], pyLBL/spectroscopy.py:15


def number_density(temperature, pressure, volume_mixing_ratio):
    """Ideal-gas number density [m-3], pyLBL/spectroscopy.py:18-29."""
    return pressure * volume_mixing_ratio / (KB * temperature)


class Mixture(object):
    """Several gases of one database on one device, summed on the device."""

    def __init__(self, lines_database, formulas, device=None, precision="fp64"):
        self.device = default_device() if device is None else int(device)
        self.gases = {f: Gas(lines_database, f, devices=[self.device], precision=precision)
                      for f in formulas}
        self._mix = None
        self._shape = None

        self._owns_gases = True
        self._auto_groups = 1
        self.last_layer_groups = 1
        self.last_copy_tail_ms = self.last_wall_ms = 0.

    @classmethod
    def from_gases(cls, gases, device=None):
        """A mixture over existing ``Gas`` objects ({formula: Gas}, all holding ``device``); they
        stay open when the mixture is closed."""
        self = cls.__new__(cls)
        self.device = default_device() if device is None else int(device)
        self.gases = dict(gases)
        self._mix = None
        self._shape = None
        self._owns_gases = False
        self._auto_groups = 1
        self.last_layer_groups = 1
        self.last_copy_tail_ms = self.last_wall_ms = 0.
        return self

    def close(self):
        if self._mix is not None:
            _lib.library().lbl_mix_close(self._mix)
            self._mix = None
        if self._owns_gases:
            for g in self.gases.values():
                g.close()
        self.gases = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # automatic layer groups: one more whenever the uncovered copy exceeds this share of a call
    TAIL_FRACTION = float(os.environ.get("PYLBL_B200_MIX_TAIL_FRACTION", "0.12"))

    @staticmethod
    def _add_stats(totals, formula, stats):
        """Statistics of a gas over the layer groups of one call: counters and times add up."""
        if formula not in totals:
            totals[formula] = dict(stats)
            return
        acc = totals[formula]
        for key in ("evals", "executed", "h2d_bytes", "d2h_bytes", "n_layers", "sum_launches",
                    "total_launches", "scale_ms", "sum_ms", "fixup_ms", "pedestal_ms", "total_ms"):
            acc[key] += stats[key]

    def total_absorption(self, temperature, pressure, volume_mixing_ratio, grid=None,
                         remove_pedestal=True, cut_off=25, bounds=None, out=None, continuum=None,
                         layer_groups=None):
        """sum over gases of n_gas * k_gas [m-1], shape (n_layers, (vn-v0)*n_per_v).

        Args:
            volume_mixing_ratio: {formula: array over layers}; every gas of the atmosphere when
                                 ``continuum`` is given (the continua read them all).
            remove_pedestal: True is what ``Spectroscopy.compute_absorption`` passes with the
                             MT-CKD continuum backend (pyLBL/spectroscopy.py:163-164).
            continuum: a ``Continuum`` on this device: the MT-CKD continua of every gas of
                       ``volume_mixing_ratio`` are added on the device too
                       (pyLBL/spectroscopy.py:194-198,225-234).
            layer_groups: the layers are taken in this many consecutive groups, all gases of a
                          group before the next group, so that a finished group's rows are on
                          their way to the host while the next group computes (only the last
                          group's copy is exposed).  Worth it when the copy is slow next to the
                          kernels (several GPUs sharing the host's PCIe and memory bandwidth);
                          every extra group costs about a millisecond of kernel tails and
                          small launches.  None = automatic: one group, and one more (up to 4)
                          whenever the previous call of this object ended with more than 12 %
                          of its time spent on a copy that nothing hid (`copy_tail_ms`).
        """
        v0, vn, n_per_v = bounds if bounds is not None else grid_to_ints(grid)
        t = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
        p = np.ascontiguousarray(pressure, dtype=np.float64).ravel()
        n_layers, n = t.size, (vn - v0) * n_per_v
        lib = _lib.library()
        if self._shape != (n_layers, n):
            if self._mix is not None:
                lib.lbl_mix_close(self._mix)
            self._mix = c_void_p()
            lib.lbl_mix_open(self.device, n_layers, n, ctypes.byref(self._mix))
            self._shape = (n_layers, n)
        else:
            lib.lbl_mix_reset(self._mix)
        if out is None:
            out = np.empty((n_layers, n))
        if out.shape != (n_layers, n) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float64 array of shape (n_layers, n)")
        if continuum is not None:
            # every continuum of every gas, summed in one pass over the accumulator
            from .continuum import continua_of
            names = [name for formula in volume_mixing_ratio for name in continua_of(formula)]
            if names:
                continuum.spectra(names, t, p, volume_mixing_ratio, bounds=(v0, vn, n_per_v), mix=self._mix)
        # Every gas is submitted without waiting for the one before: their scaling kernels and
        # pedestal chains overlap, the summation kernels run gas after gas, and each gas is
        # added into the accumulator on the device as soon as it is done.  The gas with the most
        # lines goes last in every layer group; what it completes is copied to the host while
        # the next layers compute.
        adaptive = layer_groups is None
        if adaptive:
            layer_groups = self._auto_groups if n_layers >= 8 * self._auto_groups else 1
        layer_groups = max(1, min(int(layer_groups), n_layers))
        t_begin = time.perf_counter()
        edges = [n_layers * g // layer_groups for g in range(layer_groups + 1)]
        order = sorted(self.gases.items(), key=lambda item: item[1]._handle(self.device).stats()["n_lines"])
        states = {formula: np.ascontiguousarray(volume_mixing_ratio[formula], dtype=np.float64).ravel()
                  for formula, _ in order}
        scales = {formula: np.ascontiguousarray(number_density(t, p, states[formula])) for formula, _ in order}
        handles = []
        totals = {}
        for g in range(layer_groups):
            lo, hi = edges[g], edges[g + 1]
            for i, (formula, gas) in enumerate(order):
                h = gas._handle(self.device)
                last = i == len(order) - 1
                if g > 0:
                    # (submitting on a handle waits for its previous call; its statistics are
                    # collected first)
                    lib.lbl_gas_wait(h.ptr)
                    self._add_stats(totals, formula, h.stats())
                if last:
                    # the last gas of the last group also hands its layers over in sub-groups
                    # (automatic); in the other groups that would only cost kernel tails
                    lib.lbl_gas_set_copy_groups(h.ptr, 0 if g == layer_groups - 1 else 1)
                lib.lbl_gas_submit_mix(h.ptr, hi - lo, p[lo:hi], t[lo:hi], states[formula][lo:hi],
                                       v0, vn, n_per_v, int(cut_off), 1 if remove_pedestal else 0,
                                       gas.precision, self._mix, lo, scales[formula][lo:hi],
                                       out.ctypes.data_as(c_void_p) if last else None)
                if g == 0:
                    handles.append((formula, gas, h))
        lib.lbl_mix_wait(self._mix)
        for formula, gas, h in handles:
            lib.lbl_gas_wait(h.ptr)
            self._add_stats(totals, formula, h.stats())
            gas.last_stats = [totals[formula]]
        if not handles:
            lib.lbl_mix_download(self._mix, out.ctypes.data_as(c_void_p))
        elif adaptive:
            wall_ms = (time.perf_counter() - t_begin) * 1e3
            tail_ms = handles[-1][2].stats()["copy_tail_ms"]
            self.last_copy_tail_ms, self.last_wall_ms = tail_ms, wall_ms
            if tail_ms > self.TAIL_FRACTION * wall_ms and self._auto_groups < 4:
                self._auto_groups += 1
        self.last_layer_groups = layer_groups
        return out
