"""Packed line-list cache: lbl_pack_database / lbl_pack_info need no
GPU; lbl_gas_open_pack must give spectra bit-identical to the sqlite-backed handle."""
import os

import numpy as np
import pytest

from pylbl_b200 import Gas, pack_database, pack_info, synth
from pylbl_b200.gas_optics import cached_pack


@pytest.fixture(scope="module")
def small_db(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("pack") / "small.db")
    lists = synth.config_line_lists(1, scale=0.02, seed=3)
    synth.write_database(path, lists)
    return path, lists


def test_pack_roundtrip_header(small_db, tmp_path):
    db, lists = small_db
    for formula, lines in lists.items():
        pack = pack_database(db, formula, str(tmp_path / f"{formula}.lblpack"))
        info = pack_info(pack)
        assert info["formula"] == formula
        assert info["n_lines"] == len(lines["nu"])
        assert info["sorted"] is True
        assert info["num_t"] > 0 and info["num_iso"] > 0
        st = os.stat(db)
        assert info["source_size"] == st.st_size and info["source_mtime"] == int(st.st_mtime)
        # payload = TIPS table + 8 f64 columns + i32 column (+pad) + header + checksum
        n, nt = info["n_lines"], info["num_iso"] * info["num_t"]
        payload = 16 * nt + 64 * n + 4 * (n + (n & 1))
        assert os.path.getsize(pack) > payload


def test_pack_errors(small_db, tmp_path):
    db, _ = small_db
    with pytest.raises(ValueError):
        pack_database(db, "XeF6", str(tmp_path / "nope.lblpack"))      # unknown molecule
    with pytest.raises(ValueError):
        pack_info(str(tmp_path / "missing.lblpack"))
    junk = tmp_path / "junk.lblpack"
    junk.write_bytes(b"not a pack at all" * 100)
    with pytest.raises(ValueError):
        pack_info(str(junk))


def test_cached_pack_refreshes_when_database_changes(small_db, tmp_path):
    db, lists = small_db
    cache = str(tmp_path / "cache")
    first = cached_pack(db, "H2O", cache)
    stamp = os.stat(first).st_mtime_ns
    assert cached_pack(db, "H2O", cache) == first and os.stat(first).st_mtime_ns == stamp
    os.utime(db, (os.stat(db).st_atime, os.stat(db).st_mtime + 10))   # "database was rewritten"
    cached_pack(db, "H2O", cache)
    assert pack_info(first)["source_mtime"] == int(os.stat(db).st_mtime)


def test_corrupt_pack_is_detected_without_a_gpu_and_rebuilt(small_db, tmp_path):
    """The checksum covers header and payload, and `pack_info` verifies it: a pack whose stamps
    still match the database but whose bytes are damaged is rewritten by `cached_pack`."""
    db, _ = small_db
    cache = str(tmp_path / "cache")
    pack = cached_pack(db, "CO2", cache)
    good = open(pack, "rb").read()
    for where in (40, len(good) // 2, len(good) - 3):       # header field, payload, checksum
        bad = bytearray(good)
        bad[where] ^= 0x01
        open(pack, "wb").write(bytes(bad))
        with pytest.raises(ValueError):
            pack_info(pack)
        assert cached_pack(db, "CO2", cache) == pack
        assert open(pack, "rb").read() == good


def test_concurrent_writers_leave_one_good_pack(small_db, tmp_path):
    """Several ranks sharing PYLBL_B200_CACHE write the same pack at the same time."""
    import threading
    db, _ = small_db
    target = str(tmp_path / "shared.lblpack")
    errors = []

    def work():
        try:
            for _ in range(5):
                pack_database(db, "O3", target)
        except Exception as exc:
            errors.append(exc)
    threads = [threading.Thread(target=work) for _ in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors
    assert pack_info(target)["formula"] == "O3"
    assert [f for f in os.listdir(tmp_path) if ".tmp." in f] == []


@pytest.mark.gpu
def test_pack_handle_is_bit_identical(small_db, tmp_path):
    db, _ = small_db
    atm = synth.fixture_atmosphere()
    grid = synth.grid_from_bounds(1, 501, 100)
    for formula in ("H2O", "CO2"):
        plain = Gas(db, formula)
        cached = Gas(db, formula, cache_dir=str(tmp_path / "cache"))
        assert cached.pack is not None and os.path.exists(cached.pack)
        for ped in (False, True):
            a = plain.absorption_coefficients(atm.t, atm.p, atm.vmr[formula], grid, remove_pedestal=ped)
            b = cached.absorption_coefficients(atm.t, atm.p, atm.vmr[formula], grid, remove_pedestal=ped)
            assert np.array_equal(a, b)
        plain.close()
        cached.close()


@pytest.mark.gpu
def test_truncated_pack_is_rejected(small_db, tmp_path):
    db, _ = small_db
    pack = pack_database(db, "H2O", str(tmp_path / "h2o.lblpack"))
    data = open(pack, "rb").read()
    broken = tmp_path / "broken.lblpack"
    broken.write_bytes(data[:len(data) // 2])
    from pylbl_b200.gas_optics import _Handle
    with pytest.raises(ValueError):
        _Handle(db, "H2O", 0, pack=str(broken))
    flipped = bytearray(data)
    flipped[len(data) // 2] ^= 0x40
    broken.write_bytes(bytes(flipped))
    with pytest.raises(ValueError):
        _Handle(db, "H2O", 0, pack=str(broken))
