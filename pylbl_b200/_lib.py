"""ctypes binding of libpylbl_b200.so (C ABI: include/pylbl_b200.h).

Mirrors how the reference reaches its C library: ``CDLL`` on a shared object that sits next
to the Python module (pyLBL/c_lib/gas_optics.py:11-12), ``argtypes`` per entry point
(:68-73) and a ``restype`` hook that raises ``ValueError("Error inside c functions.")`` on
a non-zero return (:15-26,76).

There is no fallback: if the shared object is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, \
    c_size_t, c_void_p
from pathlib import Path

import numpy as np
from numpy.ctypeslib import ndpointer

LIBRARY_PATH = Path(__file__).resolve().parent / "libpylbl_b200.so"

PRECISION_FP64 = 0
PRECISION_FP32 = 1


class Stats(Structure):
    """struct lbl_stats (include/pylbl_b200.h)."""
    _fields_ = [
        ("evals", c_longlong), ("executed", c_longlong), ("h2d_bytes", c_longlong),
        ("d2h_bytes", c_longlong),
        ("n_lines", c_int), ("n_active", c_int), ("n_layers", c_int), ("n_points", c_int),
        ("points_per_thread", c_int), ("cells_per_warp", c_int), ("sum_launches", c_int),
        ("total_launches", c_int), ("fp32_used", c_int),
        ("scale_ms", c_float), ("sum_ms", c_float), ("fixup_ms", c_float),
        ("pedestal_ms", c_float),
        ("total_ms", c_float), ("copy_tail_ms", c_float),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = (
    "absorption", "lbl_gas_open", "lbl_gas_close", "lbl_gas_compute", "lbl_gas_submit",
    "lbl_gas_wait", "lbl_gas_stats", "lbl_gas_device_result", "lbl_gas_windows",
    "lbl_gas_scaled", "lbl_host_alloc", "lbl_host_free", "lbl_device_count",
    "lbl_set_chunk_layers", "lbl_last_error", "lbl_version", "lbl_timer_start",
    "lbl_timer_join", "lbl_timer_stop", "lbl_measure_fp64_peak", "lbl_mix_open", "lbl_mix_reset",
    "lbl_mix_add", "lbl_mix_download", "lbl_mix_close", "lbl_pack_database", "lbl_pack_info",
    "lbl_gas_open_pack", "lbl_gas_set_copy_groups", "lbl_gas_submit_band", "lbl_gas_submit_mix",
    "lbl_mix_wait", "lbl_mix_device_result", "lbl_gas_band_edges", "lbl_continuum_create",
    "lbl_continuum_set_spectrum", "lbl_continuum_finalize", "lbl_continuum_compute",
    "lbl_continuum_close", "lbl_continuum_last_ms",
)

_library = None

# The library runs a handful of CUDA streams per device (pylbl_b200/csrc/lbl_api.cu,
# DeviceStreams); the default of 8 hardware work queues per context makes some of them share a
# queue, and work then waits for unrelated work.  Only effective before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def check_return_code(value):
    """Same contract as pyLBL/c_lib/gas_optics.py:15-26."""
    if value != 0:
        raise ValueError("Error inside c functions.")
    return value


def library():
    """Loads the CUDA library (once) and declares its signatures."""
    global _library
    if _library is not None:
        return _library
    if not LIBRARY_PATH.exists():
        raise ImportError(
            f"{LIBRARY_PATH} is missing: build it with `python -m pylbl_b200.build` "
            "(this backend has no CPU path)")
    lib = ctypes.CDLL(str(LIBRARY_PATH))
    f64 = ndpointer(np.float64, flags="C_CONTIGUOUS")
    i32 = ndpointer(np.int32, flags="C_CONTIGUOUS")

    # The reference's own entry point, argtypes exactly as pyLBL/c_lib/gas_optics.py:68-73.
    lib.absorption.argtypes = 3 * [c_double] + 3 * [c_int] + [f64] + 2 * [c_char_p] + 2 * [c_int]
    lib.absorption.restype = check_return_code

    lib.lbl_gas_open.argtypes = [c_char_p, c_char_p, c_int, POINTER(c_void_p)]
    lib.lbl_gas_close.argtypes = [c_void_p]
    lib.lbl_pack_database.argtypes = [c_char_p, c_char_p, c_char_p]
    lib.lbl_pack_info.argtypes = [c_char_p, c_char_p, c_int, POINTER(c_longlong), POINTER(c_int),
                                  POINTER(c_int), POINTER(c_int), POINTER(c_longlong),
                                  POINTER(c_longlong)]
    lib.lbl_gas_open_pack.argtypes = [c_char_p, c_int, POINTER(c_void_p)]
    batched = [c_void_p, c_int, f64, f64, f64] + 6 * [c_int] + [c_void_p]
    lib.lbl_gas_compute.argtypes = batched
    lib.lbl_gas_submit.argtypes = batched
    lib.lbl_gas_submit_band.argtypes = [c_void_p, c_int, f64, f64, f64] + 6 * [c_int] + \
        [c_int, c_int, c_void_p, c_longlong]
    lib.lbl_gas_band_edges.argtypes = [c_void_p] + 5 * [c_int] + [i32]
    lib.lbl_gas_submit_mix.argtypes = [c_void_p, c_int, f64, f64, f64] + 6 * [c_int] + \
        [c_void_p, c_int, f64, c_void_p]
    lib.lbl_continuum_create.argtypes = [c_int, POINTER(c_void_p)]
    lib.lbl_continuum_set_spectrum.argtypes = [c_void_p, c_char_p, c_double, c_double, c_double, c_int, f64]
    lib.lbl_continuum_finalize.argtypes = [c_void_p]
    lib.lbl_continuum_compute.argtypes = [c_void_p, c_char_p, c_int, f64, f64, f64, c_int, c_int, c_int,
                                          c_void_p, c_int, c_void_p]
    lib.lbl_continuum_close.argtypes = [c_void_p]
    lib.lbl_continuum_last_ms.argtypes = [c_void_p, POINTER(c_float), POINTER(c_float)]
    lib.lbl_mix_wait.argtypes = [c_void_p]
    lib.lbl_mix_device_result.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_longlong)]
    lib.lbl_gas_wait.argtypes = [c_void_p]
    lib.lbl_gas_set_copy_groups.argtypes = [c_void_p, c_int]
    lib.lbl_gas_stats.argtypes = [c_void_p, POINTER(Stats)]
    lib.lbl_gas_device_result.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_longlong)]
    lib.lbl_gas_windows.argtypes = [c_void_p, c_int, i32, i32, c_int]
    lib.lbl_gas_scaled.argtypes = [c_void_p, c_int, f64, c_int]
    lib.lbl_host_alloc.argtypes = [c_size_t, POINTER(c_void_p)]
    lib.lbl_host_free.argtypes = [c_void_p]
    lib.lbl_device_count.argtypes = [POINTER(c_int)]
    lib.lbl_set_chunk_layers.argtypes = [c_int]
    lib.lbl_timer_start.argtypes = [c_int]
    lib.lbl_timer_join.argtypes = [c_void_p]
    lib.lbl_timer_stop.argtypes = [c_int, POINTER(c_float)]
    lib.lbl_measure_fp64_peak.argtypes = [c_int, POINTER(c_double)]
    lib.lbl_mix_open.argtypes = [c_int, c_int, c_int, POINTER(c_void_p)]
    lib.lbl_mix_reset.argtypes = [c_void_p]
    lib.lbl_mix_add.argtypes = [c_void_p, c_void_p, f64]
    lib.lbl_mix_download.argtypes = [c_void_p, c_void_p]
    lib.lbl_mix_close.argtypes = [c_void_p]
    for name in EXPORTS:
        if name not in ("absorption", "lbl_last_error", "lbl_version"):
            getattr(lib, name).restype = check_return_code
    lib.lbl_last_error.argtypes = []
    lib.lbl_last_error.restype = c_char_p
    lib.lbl_version.argtypes = []
    lib.lbl_version.restype = c_int
    _library = lib
    return lib


def last_error() -> str:
    return library().lbl_last_error().decode("utf-8", "replace")


def device_count() -> int:
    n = c_int(0)
    try:
        library().lbl_device_count(ctypes.byref(n))
    except ValueError:
        return 0
    return int(n.value)


class PinnedArray(object):
    """A float64 numpy array over page-locked host memory from lbl_host_alloc()."""

    def __init__(self, shape):
        count = int(np.prod(shape))
        self._ptr = c_void_p()
        library().lbl_host_alloc(max(count, 1) * 8, ctypes.byref(self._ptr))
        buf = (c_double * max(count, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=count).reshape(shape)

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr is not None and ptr.value and _library is not None:
            try:
                _library.lbl_host_free(ptr)
            except Exception:
                pass
