"""Shared fixtures.  GPU tests are marked ``gpu``; everything else runs on CPU."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from pylbl_b200 import synth  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        from pylbl_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_db(tmp_path_factory):
    """H2O/CO2/O3, 2000+2000+1000 lines over 0.5-5025 cm-1 (config 1 at 1/10 scale)."""
    path = tmp_path_factory.mktemp("db") / "config1_small.db"
    synth.write_database(str(path), synth.config_line_lists(1, scale=0.1))
    return str(path)


@pytest.fixture(scope="session")
def dense_db(tmp_path_factory):
    """CO2 only, 1500 lines inside 474.5-875.5 cm-1 (config 3 at 1/40 scale)."""
    path = tmp_path_factory.mktemp("db") / "config3_small.db"
    synth.write_database(str(path), synth.config_line_lists(3, scale=0.025))
    return str(path)


@pytest.fixture(scope="session")
def atmosphere():
    """The reference's 4-layer fixture atmosphere (tests/conftest.py:54-78 there)."""
    return synth.fixture_atmosphere()
