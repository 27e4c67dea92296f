"""The B200 lines backend: a ``Gas`` class with pyLBL's plugin signature.

Host-side mirror of the reference adapter pyLBL/c_lib/gas_optics.py:29-92: same
constructor ``Gas(lines_database, formula)``, same method
``absorption_coefficient(temperature, pressure, volume_mixing_ratio, grid,
remove_pedestal=False, cut_off=25)``, same return value (float64 array of
``(vn - v0)*n_per_v`` cross-sections [m2 per molecule], pyLBL/c_lib/gas_optics.py:61-65,92)
and the same error (``ValueError("Error inside c functions.")``, :15-26).

What differs is behind the boundary: the sqlite database is read once per
(database, formula) and packed into device memory; a call computes one layer -- or, through
``absorption_coefficients`` / ``prefetch``, a whole column of layers -- on the GPU.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_int, c_longlong, c_void_p

import numpy as np

from . import _lib


def grid_to_ints(grid):
    """(v0, vn, n_per_v) exactly as pyLBL/c_lib/gas_optics.py:61-63 derives them."""
    v0 = int(round(grid[0]))
    vn = int(round(grid[-1]) + 1)
    n_per_v = int(round(1. / (grid[1] - grid[0])))
    return v0, vn, n_per_v


def default_device() -> int:
    """PYLBL_B200_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    for name in ("PYLBL_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(name, "") != "":
            return int(os.environ[name])
    return 0


def pack_info(pack_path):
    """Header of a packed line-list file (no GPU needed)."""
    formula = ctypes.create_string_buffer(64)
    n_lines, size, mtime = c_longlong(0), c_longlong(0), c_longlong(0)
    num_iso, num_t, is_sorted = c_int(0), c_int(0), c_int(0)
    _lib.library().lbl_pack_info(os.fsencode(pack_path), formula, 64, ctypes.byref(n_lines),
                                 ctypes.byref(num_iso), ctypes.byref(num_t),
                                 ctypes.byref(is_sorted), ctypes.byref(size), ctypes.byref(mtime))
    return {"formula": formula.value.decode(), "n_lines": n_lines.value, "num_iso": num_iso.value,
            "num_t": num_t.value, "sorted": bool(is_sorted.value), "source_size": size.value,
            "source_mtime": mtime.value}


def pack_database(database, formula, pack_path):
    """Writes the packed line-list cache of one molecule (no GPU needed): what the reference
    re-reads from sqlite on every call (pyLBL/c_lib/absorption.c:45-79), once, as one flat
    binary file.  Returns ``pack_path``."""
    _lib.library().lbl_pack_database(os.fsencode(getattr(database, "path", database)),
                                     formula.encode("utf-8"), os.fsencode(pack_path))
    return pack_path


def cached_pack(database, formula, cache_dir):
    """Path of the pack of (database, formula) under ``cache_dir``, written or rewritten when
    it is missing, unreadable, or older than the sqlite file."""
    database = getattr(database, "path", database)
    os.makedirs(cache_dir, exist_ok=True)
    stem = os.path.basename(database)
    path = os.path.join(cache_dir, f"{stem}.{formula}.lblpack")
    st = os.stat(database)
    try:
        info = pack_info(path) if os.path.exists(path) else None
    except ValueError:
        info = None
    if (info is None or info["formula"] != formula or info["source_size"] != st.st_size
            or info["source_mtime"] != int(st.st_mtime)):
        pack_database(database, formula, path)
    return path


class _Handle(object):
    """Owns one lbl_gas* (one molecule packed on one device)."""

    def __init__(self, database, formula, device, pack=None):
        self.ptr = c_void_p()
        self.device = device
        if pack is not None:
            _lib.library().lbl_gas_open_pack(os.fsencode(pack), int(device), ctypes.byref(self.ptr))
        else:
            _lib.library().lbl_gas_open(bytes(database, encoding="utf-8"),
                                        bytes(formula, encoding="utf-8"), int(device),
                                        ctypes.byref(self.ptr))

    def close(self):
        ptr, self.ptr = self.ptr, None
        if ptr is not None and ptr.value:
            _lib.library().lbl_gas_close(ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = _lib.Stats()
        _lib.library().lbl_gas_stats(self.ptr, ctypes.byref(s))
        return s.as_dict()


class Gas(object):
    """API for gas optics calculation on B200 GPUs.

    Attributes:
        database: String path to the spectral sqlite3 database.
        formula: String chemical formula.
        devices: CUDA device indices this object spreads a call over.
        shard: "layer" (contiguous layer shards, one per device) or "band" (contiguous spectral
               bands of about equal work, one per device; every device computes all layers).
        precision: "fp64" (default) or "fp32".
    """

    def __init__(self, lines_database, formula, devices=None, precision="fp64", cache_dir=None,
                 shard="layer"):
        """Initializes the object.

        Args:
            lines_database: Database object (only its ``path`` is read,
                            pyLBL/c_lib/gas_optics.py:43) or a path string.
            formula: String chemical formula.
            devices: int, list of ints, or None (see default_device()).
            shard: how a call is split over several devices, "layer" or "band" (SURVEY.md 8(e):
                   layers/columns for many states, spectral bands for one huge spectrum).
            precision: "fp64" or "fp32".
            cache_dir: directory for the packed line-list cache (default: the environment
                       variable PYLBL_B200_CACHE, else no cache).  With a cache the sqlite file
                       is read once per (database, formula) ever, not once per process.
        """
        self.database = getattr(lines_database, "path", lines_database)
        self.formula = formula
        if devices is None:
            devices = [default_device()]
        elif isinstance(devices, int):
            devices = [devices]
        self.devices = list(devices)
        if shard not in ("layer", "band"):
            raise ValueError("shard must be 'layer' or 'band'")
        self.shard = shard
        self.precision = {"fp64": _lib.PRECISION_FP64, "fp32": _lib.PRECISION_FP32}[precision]
        self._handles = {}
        self._cache = {}
        self.last_stats = []
        if cache_dir is None:
            cache_dir = os.environ.get("PYLBL_B200_CACHE") or None
        self.cache_dir = cache_dir
        self.pack = cached_pack(self.database, formula, cache_dir) if cache_dir else None

    # -- handles -----------------------------------------------------------------------
    def _handle(self, device):
        if device not in self._handles:
            try:
                self._handles[device] = _Handle(self.database, self.formula, device, pack=self.pack)
            except ValueError:
                if self.pack is None or not os.path.exists(self.database):
                    raise
                # A pack that passed the header check but does not open (truncated or corrupt
                # payload): read the sqlite file instead and write the pack anew.
                self._handles[device] = _Handle(self.database, self.formula, device, pack=None)
                try:
                    pack_database(self.database, self.formula, self.pack)
                except ValueError:
                    self.pack = None
        return self._handles[device]

    def band_edges(self, bounds, n_bands, cut_off=25):
        """Cell indices edges[0..n_bands] of ``n_bands`` contiguous spectral bands of about equal
        work on the grid ``bounds = (v0, vn, n_per_v)``; band b is cells [edges[b], edges[b+1])."""
        v0, vn, n_per_v = bounds
        edges = np.zeros(n_bands + 1, dtype=np.int32)
        _lib.library().lbl_gas_band_edges(self._handle(self.devices[0]).ptr, v0, vn, n_per_v,
                                          int(cut_off), int(n_bands), edges)
        return edges

    def close(self):
        for h in self._handles.values():
            h.close()
        self._handles = {}
        self._cache = {}

    # -- the plugin method ---------------------------------------------------------------
    def absorption_coefficient(self, temperature, pressure, volume_mixing_ratio, grid,
                               remove_pedestal=False, cut_off=25):
        """Calculates absorption coefficient.

        Args:
            temperature: Temperature [K].
            pressure: Pressure [Pa].
            volume_mixing_ratio: Volume mixing ratio [mol mol-1].
            grid: Numpy array defining the spectral grid [cm-1].
            remove_pedestal: Flag specifying if a pedestal should be subtracted.
            cut_off: Wavenumber cut-off distance [cm-1] from line centers.

        Returns:
            Numpy array of absorption coefficients [m2].
        """
        v0, vn, n_per_v = grid_to_ints(grid)
        key = (float(temperature), float(pressure), float(volume_mixing_ratio), v0, vn, n_per_v,
               int(cut_off), bool(remove_pedestal))
        row = self._cache.pop(key, None)
        if row is not None:
            return row
        k = self.absorption_coefficients([temperature], [pressure], [volume_mixing_ratio], grid,
                                         remove_pedestal=remove_pedestal, cut_off=cut_off)
        return k[0]

    # -- batched forms -------------------------------------------------------------------
    def absorption_coefficients(self, temperature, pressure, volume_mixing_ratio, grid=None,
                                remove_pedestal=False, cut_off=25, bounds=None, out=None,
                                to_host=True):
        """Layer-batched ``absorption_coefficient``: row L of the result is what the scalar
        call returns for (temperature[L], pressure[L], volume_mixing_ratio[L]).

        Args:
            grid: spectral grid array, or None with ``bounds=(v0, vn, n_per_v)``.
            out: optional C-contiguous float64 array (n_layers, n) to fill (e.g. pinned).
            to_host: False leaves the spectra on the device (benchmarks) and returns None.

        With several devices the call is split into contiguous layer shards
        (``shard="layer"``) or into spectral bands of the grid (``shard="band"``, see
        ``absorption_band``), one per device; the pieces are independent, so there is no
        inter-device communication.
        """
        v0, vn, n_per_v = bounds if bounds is not None else grid_to_ints(grid)
        t = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
        p = np.ascontiguousarray(pressure, dtype=np.float64).ravel()
        x = np.ascontiguousarray(volume_mixing_ratio, dtype=np.float64).ravel()
        n_layers = t.size
        n = (vn - v0) * n_per_v
        if to_host and out is None:
            out = np.empty((n_layers, n))
        if to_host and (out.shape != (n_layers, n) or out.dtype != np.float64
                        or not out.flags["C_CONTIGUOUS"]):
            raise ValueError("out must be a C-contiguous float64 array of shape (n_layers, n)")
        lib = _lib.library()
        ped = 1 if remove_pedestal else 0
        if self.shard == "band" and len(self.devices) > 1 and to_host:
            return self._band_sharded(t, p, x, (v0, vn, n_per_v), ped, int(cut_off), out)
        ndev = max(1, min(len(self.devices), n_layers))
        edges = np.linspace(0, n_layers, ndev + 1).astype(int)
        submitted = []
        errors = []

        def submit(d, blocking=False):
            lo, hi = int(edges[d]), int(edges[d + 1])
            h = self._handle(self.devices[d])
            dst = out[lo:hi].ctypes.data_as(c_void_p) if to_host else None
            # The blocking entry point copies the spectra out in layer groups while later groups
            # compute; submit + wait leaves that overlap to the calls queued behind this one.
            call = lib.lbl_gas_compute if blocking else lib.lbl_gas_submit
            call(h.ptr, hi - lo, p[lo:hi], t[lo:hi], x[lo:hi], v0, vn, n_per_v,
                 int(cut_off), ped, self.precision, dst)
            return h

        if ndev == 1:
            submitted.append(submit(0, blocking=True))
        else:
            # One host thread per device: pageable destinations make the copies synchronous.
            def work(d):
                try:
                    submitted.append(submit(d, blocking=True))
                except Exception as exc:  # re-raised below
                    errors.append(exc)
            threads = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            if errors:
                raise errors[0]
        for h in submitted:
            lib.lbl_gas_wait(h.ptr)
        self.last_stats = [h.stats() for h in submitted]
        return out if to_host else None

    def submit(self, temperature, pressure, volume_mixing_ratio, grid=None, remove_pedestal=False,
               cut_off=25, bounds=None, out=None, device_index=0):
        """Non-blocking ``absorption_coefficients`` on one device: the work is enqueued and the
        call returns; ``wait()`` completes it.  Several gases submitted back to back overlap on
        the device (scaling kernels and pedestal chains side by side, summation kernels gas after
        gas, each gas's copy to the host under the next gas's kernels).  ``out``: C-contiguous
        float64 (n_layers, n) -- page-locked (``_lib.PinnedArray``) for the copy to be
        asynchronous -- or None to leave the spectra on the device."""
        v0, vn, n_per_v = bounds if bounds is not None else grid_to_ints(grid)
        t = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
        p = np.ascontiguousarray(pressure, dtype=np.float64).ravel()
        x = np.ascontiguousarray(volume_mixing_ratio, dtype=np.float64).ravel()
        n = (vn - v0) * n_per_v
        if out is not None and (out.shape != (t.size, n) or out.dtype != np.float64
                                or not out.flags["C_CONTIGUOUS"]):
            raise ValueError("out must be a C-contiguous float64 array of shape (n_layers, n)")
        h = self._handle(self.devices[device_index])
        _lib.library().lbl_gas_submit(h.ptr, t.size, p, t, x, v0, vn, n_per_v, int(cut_off),
                                      1 if remove_pedestal else 0, self.precision,
                                      out.ctypes.data_as(c_void_p) if out is not None else None)
        self._submitted = h
        return out

    def wait(self):
        """Completes the last ``submit``; returns its statistics."""
        h = getattr(self, "_submitted", None)
        if h is None:
            return None
        _lib.library().lbl_gas_wait(h.ptr)
        self._submitted = None
        self.last_stats = [h.stats()]
        return self.last_stats[0]

    def absorption_band(self, temperature, pressure, volume_mixing_ratio, bounds, band,
                        remove_pedestal=False, cut_off=25, out=None, device_index=0, wait=True):
        """Cells ``band = (lo, hi)`` of the grid ``bounds``: columns [lo*n_per_v, hi*n_per_v) of
        what ``absorption_coefficients`` returns for the whole grid, bit for bit (the line
        windows, the active-line prefix and the pedestal are those of the whole grid).  ``out``
        may be a column slice of a whole-grid array (rows C-contiguous)."""
        v0, vn, n_per_v = bounds
        lo, hi = int(band[0]), int(band[1])
        t = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
        p = np.ascontiguousarray(pressure, dtype=np.float64).ravel()
        x = np.ascontiguousarray(volume_mixing_ratio, dtype=np.float64).ravel()
        m = (hi - lo) * n_per_v
        if out is None:
            out = np.empty((t.size, m))
        if out.shape != (t.size, m) or out.dtype != np.float64 or out.strides[1] != 8 \
                or out.strides[0] % 8 or out.strides[0] < 8 * m:
            raise ValueError("out must be float64 of shape (n_layers, band points), rows contiguous")
        h = self._handle(self.devices[device_index])
        _lib.library().lbl_gas_submit_band(h.ptr, t.size, p, t, x, v0, vn, n_per_v, int(cut_off),
                                           1 if remove_pedestal else 0, self.precision, lo, hi,
                                           out.ctypes.data_as(c_void_p), out.strides[0] // 8)
        if wait:
            _lib.library().lbl_gas_wait(h.ptr)
            self.last_stats = [h.stats()]
        return out

    def _band_sharded(self, t, p, x, bounds, ped, cut_off, out):
        """One band per device (``shard="band"``), every band written into its columns of ``out``."""
        n_per_v = bounds[2]
        ndev = len(self.devices)
        edges = self.band_edges(bounds, ndev, cut_off)
        errors, handles = [], []

        def work(d):
            lo, hi = int(edges[d]), int(edges[d + 1])
            if hi <= lo:
                return
            try:
                self.absorption_band(t, p, x, bounds, (lo, hi), remove_pedestal=bool(ped),
                                     cut_off=cut_off, out=out[:, lo * n_per_v:hi * n_per_v],
                                     device_index=d, wait=False)
                h = self._handle(self.devices[d])
                _lib.library().lbl_gas_wait(h.ptr)
                handles.append((d, h))
            except Exception as exc:  # re-raised below
                errors.append(exc)
        threads = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        self.last_stats = [h.stats() for _, h in sorted(handles, key=lambda e: e[0])]
        return out

    def prefetch(self, temperature, pressure, volume_mixing_ratio, grid, remove_pedestal=False,
                 cut_off=25):
        """Computes every layer of a column in one batch and keeps the rows, so that the
        scalar calls pyLBL's driver loop makes next (pyLBL/spectroscopy.py:179-191) return
        immediately.  Rows are handed out once and dropped."""
        v0, vn, n_per_v = grid_to_ints(grid)
        k = self.absorption_coefficients(temperature, pressure, volume_mixing_ratio, grid,
                                         remove_pedestal=remove_pedestal, cut_off=cut_off)
        t = np.asarray(temperature, dtype=np.float64).ravel()
        p = np.asarray(pressure, dtype=np.float64).ravel()
        x = np.asarray(volume_mixing_ratio, dtype=np.float64).ravel()
        for i in range(t.size):
            key = (float(t[i]), float(p[i]), float(x[i]), v0, vn, n_per_v, int(cut_off),
                   bool(remove_pedestal))
            self._cache[key] = k[i]
        return k

    # -- introspection used by the parity tests -------------------------------------------
    def windows(self, layer=0, device_index=0):
        """(s, e) of pyLBL/c_lib/spectra.c:48-62 per processed line of the last call."""
        h = self._handle(self.devices[device_index])
        n = h.stats()["n_active"]
        s = np.full(max(n, 1), -1, dtype=np.int32)
        e = np.full(max(n, 1), -1, dtype=np.int32)
        _lib.library().lbl_gas_windows(h.ptr, int(layer), s, e, int(s.size))
        return np.stack([s[:n], e[:n]], axis=1)

    def scaled_lines(self, layer=0, device_index=0):
        """(nu', alpha, gamma, sw') of pyLBL/c_lib/spectra.c:17-45 per processed line."""
        h = self._handle(self.devices[device_index])
        n = h.stats()["n_active"]
        out = np.zeros((max(n, 1), 4))
        _lib.library().lbl_gas_scaled(h.ptr, int(layer), out, int(out.shape[0]))
        return out[:n]

    def device_result(self, device_index=0):
        """(device pointer, element count) of the spectra left on the device."""
        h = self._handle(self.devices[device_index])
        ptr = c_void_p()
        count = c_longlong(0)
        _lib.library().lbl_gas_device_result(h.ptr, ctypes.byref(ptr), ctypes.byref(count))
        return ptr.value, int(count.value)
