"""The C-ABI library loads, exports every symbol include/pylbl_b200.h declares, and the
host-side plugin mirror behaves like the reference adapter -- without any GPU compute."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import pylbl_b200
from pylbl_b200 import Gas, _lib, grid_to_ints, synth

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pylbl_b200.h").read_text()
    return re.findall(r"LBL_API\s+[A-Za-z_ \*]+?\b(absorption|lbl_[a-z0-9_]+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert "absorption" in names and len(names) >= 15
    lib = ctypes.CDLL(str(_lib.LIBRARY_PATH))
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(names) == sorted(_lib.EXPORTS)
    assert _lib.library().lbl_version() >= 100


def test_grid_to_ints_matches_reference_rule():
    # pyLBL/c_lib/gas_optics.py:61-63 on the reference's own fixtures.
    assert grid_to_ints(np.arange(1., 3250., 0.1)) == (1, 3251, 10)
    assert grid_to_ints(np.arange(1., 3000., 1.)) == (1, 3000, 1)
    assert grid_to_ints(np.arange(1., 5000., 0.01)) == (1, 5001, 100)
    assert grid_to_ints(synth.grid_from_bounds(500, 851, 2000)) == (500, 851, 2000)


def test_registration_is_guarded():
    # pyLBL is not importable in this image (xarray/sqlalchemy absent): registration must
    # report False rather than raise.  When pyLBL is installed it returns True.
    try:
        import pyLBL  # noqa: F401
        assert pylbl_b200.registered is True
    except Exception:
        assert pylbl_b200.registered is False
    assert pylbl_b200.BACKEND_NAME == "b200"


def test_constructor_reads_only_path(small_db):
    class Database(object):
        path = small_db

        def __getattr__(self, name):
            raise AssertionError(f"backend touched Database.{name}")
    gas = Gas(Database(), "H2O")
    assert gas.database == small_db and gas.formula == "H2O"


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(small_db):
    """No CPU fallback: on a machine without a CUDA device the call raises the reference's
    own error type and the library says why."""
    gas = Gas(small_db, "H2O")
    with pytest.raises(ValueError, match="Error inside c functions."):
        gas.absorption_coefficient(288.99, 98388., 6.6e-3, np.arange(1., 50., 0.1))
    assert "CUDA" in _lib.last_error() or "no CUDA device" in _lib.last_error()
    k = np.zeros(490)
    with pytest.raises(ValueError):
        _lib.library().absorption(98388., 288.99, 6.6e-3, 1, 50, 10, k,
                                  bytes(small_db, "utf-8"), b"H2O", 25, 0)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "pylbl_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            text = path.read_text()
            assert "import oracle" not in text and "from oracle" not in text, path
            assert "lbl_oracle" not in text, path


def test_synthetic_inputs_are_deterministic(tmp_path):
    a = synth.make_line_list("CO2", 500, 0.5, 900.0, seed=4)
    b = synth.make_line_list("CO2", 500, 0.5, 900.0, seed=4)
    for key in a:
        assert np.array_equal(a[key], b[key])
    assert np.all(np.diff(a["nu"]) >= 0)
    col = synth.standard_column(60)
    assert col.p.size == 60 and np.all(np.diff(col.p) < 0)
    assert 150.0 < col.t.min() and col.t.max() < 330.0
    for c in range(1, 5):
        pert = synth.standard_column(60, column=c)
        assert 150.0 < pert.t.min() and pert.t.max() < 330.0
