"""Parity of the CUDA path against the oracle, through the C ABI (ctypes).

Mirrors the reference's tests/test_gas_optics.py: same fixture atmosphere, same default
keyword arguments, `Gas(database, formula).absorption_coefficient(...)`.
"""
import numpy as np
import pytest

from oracle import OracleGas, oracle_scale_line, oracle_tips
from pylbl_b200 import Gas, synth
from pylbl_b200 import _lib

from helpers import FP32_TOL, FP64_TOL, relative_error, scaled_error

pytestmark = pytest.mark.gpu


class Db(object):
    """Stand-in for pyLBL's Database: the backend only reads ``.path``."""
    def __init__(self, path):
        self.path = path


@pytest.mark.parametrize("n_per_v,farfield", [(10, None), (100, None), (100, "2"), (100, "0")])
@pytest.mark.parametrize("formula", ["H2O", "CO2", "O3"])
@pytest.mark.parametrize("remove_pedestal", [False, True])
def test_fixture_atmosphere(small_db, atmosphere, formula, n_per_v, farfield, remove_pedestal,
                            monkeypatch):
    """The reference's own test shape (tests/test_gas_optics.py:8-19).  On the fine grid the
    summation kernel is also forced either way (PYLBL_B200_FARFIELD=2: far-field kernel K2c,
    =0: direct kernel K2), so that both are checked against the oracle whatever the selector
    would pick for these short line lists."""
    if farfield is not None:
        monkeypatch.setenv("PYLBL_B200_FARFIELD", farfield)
    grid = synth.grid_from_bounds(1, 1201, n_per_v)
    gas = Gas(Db(small_db), formula)
    ref = OracleGas(Db(small_db), formula)
    for layer in range(atmosphere.t.size):
        args = (atmosphere.t[layer], atmosphere.p[layer], atmosphere.vmr[formula][layer], grid)
        k = gas.absorption_coefficient(*args, remove_pedestal=remove_pedestal)
        k_ref = ref.absorption_coefficient(*args, remove_pedestal=remove_pedestal)
        assert k.shape == k_ref.shape and k.dtype == np.float64
        assert scaled_error(k, k_ref, n_per_v) <= FP64_TOL
        if not remove_pedestal:
            assert relative_error(k, k_ref) <= FP64_TOL
        assert gas.last_stats[0]["evals"] == ref.last_evals
        if farfield is not None:
            assert (gas.last_stats[0]["cells_per_warp"] > 0) == (farfield == "2")


def test_windows_bit_exact(small_db, atmosphere):
    for formula in ("H2O", "CO2", "O3"):
        gas = Gas(small_db, formula)
        ref = OracleGas(small_db, formula)
        for n_per_v, v0, vn in ((10, 1, 5001), (100, 200, 900), (3, 1, 3000)):
            gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr[formula],
                                        bounds=(v0, vn, n_per_v))
            for layer in range(atmosphere.t.size):
                ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                               atmosphere.vmr[formula][layer], v0, vn, n_per_v, windows=True)
                win = gas.windows(layer)
                assert win.shape[0] == ref.last_active
                assert np.array_equal(win, ref.last_windows[:ref.last_active])


def test_scaled_line_parameters(small_db, atmosphere):
    formula = "CO2"
    gas = Gas(small_db, formula)
    ref = OracleGas(small_db, formula)
    d = ref.data
    gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr[formula],
                                bounds=(1, 5001, 10))
    for layer in range(atmosphere.t.size):
        got = gas.scaled_lines(layer)
        t, p, x = atmosphere.t[layer], atmosphere.p[layer], atmosphere.vmr[formula][layer]
        want = np.zeros_like(got)
        for r in range(got.shape[0]):
            iso = int(d["local_iso_id"][r])
            q_ref = oracle_tips(d["tips_t"], d["tips_q"], d["num_t"], 296., iso - 1)
            q_t = oracle_tips(d["tips_t"], d["tips_q"], d["num_t"], t, iso - 1)
            want[r] = oracle_scale_line(t, p, x, d["nu"][r], d["sw"][r], d["gamma_air"][r],
                                        d["gamma_self"][r], d["n_air"][r], d["elower"][r],
                                        d["delta_air"][r], d["mass"][iso - 1], q_ref, q_t)
        assert np.array_equal(got[:, 0], want[:, 0])          # shifted centre: bit-exact
        np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=1e-12, atol=0)


def test_reference_entry_point(small_db, atmosphere):
    """The library exports the reference's own symbol with its own signature
    (pyLBL/c_lib/absorption.c:19-30), callable exactly as gas_optics.py:79-91 does."""
    lib = _lib.library()
    v0, vn, n_per_v = 1, 801, 10
    k = np.zeros((vn - v0) * n_per_v)
    layer = 3
    lib.absorption(float(atmosphere.p[layer]), float(atmosphere.t[layer]),
                   float(atmosphere.vmr["H2O"][layer]), v0, vn, n_per_v, k,
                   bytes(small_db, encoding="utf-8"), b"H2O", 25, 0)
    k_ref = OracleGas(small_db, "H2O").absorption(atmosphere.t[layer], atmosphere.p[layer],
                                                  atmosphere.vmr["H2O"][layer], v0, vn, n_per_v)
    assert relative_error(k, k_ref) <= FP64_TOL


def test_batched_equals_scalar_and_prefetch(small_db, atmosphere):
    grid = synth.grid_from_bounds(1, 601, 100)
    gas = Gas(small_db, "O3")
    batch = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["O3"], grid,
                                        remove_pedestal=True)
    for layer in range(atmosphere.t.size):
        one = gas.absorption_coefficient(atmosphere.t[layer], atmosphere.p[layer],
                                         atmosphere.vmr["O3"][layer], grid, remove_pedestal=True)
        assert np.array_equal(one, batch[layer])
    gas.prefetch(atmosphere.t, atmosphere.p, atmosphere.vmr["O3"], grid, remove_pedestal=True)
    assert len(gas._cache) == atmosphere.t.size
    one = gas.absorption_coefficient(atmosphere.t[2], atmosphere.p[2], atmosphere.vmr["O3"][2],
                                     grid, remove_pedestal=True)
    assert np.array_equal(one, batch[2]) and len(gas._cache) == atmosphere.t.size - 1


@pytest.mark.parametrize("bounds", [(1, 3000, 1), (1, 400, 3), (1, 300, 7), (1, 200, 8),
                                    (1, 120, 1000), (640, 700, 2000), (1, 5001, 4)])
@pytest.mark.parametrize("remove_pedestal", [False, True])
def test_grid_shapes(small_db, atmosphere, bounds, remove_pedestal):
    """Every points-per-thread specialisation (n_per_v = 1, 3, 7 -> 1; 8; 1000, 2000 -> 10;
    4) and grids that start inside the line list (the early break then yields zeros,
    absorption.c:80-83)."""
    v0, vn, n_per_v = bounds
    gas = Gas(small_db, "CO2")
    ref = OracleGas(small_db, "CO2")
    k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"],
                                    bounds=bounds, remove_pedestal=remove_pedestal)
    for layer in range(atmosphere.t.size):
        k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                               atmosphere.vmr["CO2"][layer], v0, vn, n_per_v, remove_pedestal)
        if not np.any(k_ref):
            assert not np.any(k[layer])
            continue
        assert scaled_error(k[layer], k_ref, n_per_v) <= FP64_TOL


@pytest.mark.parametrize("n_per_v", [64, 65, 96, 127, 129, 250, 333])
def test_far_field_kernel_on_odd_grids(small_db, atmosphere, n_per_v, monkeypatch):
    """The far-field kernel where the points of a cell do not fill its lanes evenly (the direct
    part takes 128 points per pass, the interpolation 32 per step): points per cm-1 at and
    around the kernel's lower limit and the pass sizes, forced whatever the line density."""
    monkeypatch.setenv("PYLBL_B200_FARFIELD", "2")
    bounds = (1, 251, n_per_v)
    gas = Gas(small_db, "H2O")
    ref = OracleGas(small_db, "H2O")
    for remove_pedestal in (False, True):
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["H2O"],
                                        bounds=bounds, remove_pedestal=remove_pedestal)
        assert gas.last_stats[0]["cells_per_warp"] == 1
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["H2O"][layer], *bounds, remove_pedestal)
            assert np.any(k_ref)
            if remove_pedestal:
                assert scaled_error(k[layer], k_ref, n_per_v) <= FP64_TOL
            else:
                assert relative_error(k[layer], k_ref) <= FP64_TOL
    gas.close()


@pytest.mark.parametrize("cut_off", [0, 1, 5, 40, 70])
def test_cut_off(small_db, atmosphere, cut_off):
    gas = Gas(small_db, "H2O")
    ref = OracleGas(small_db, "H2O")
    bounds = (1, 500, 10)
    for ped in (False, True):
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["H2O"],
                                        bounds=bounds, remove_pedestal=ped, cut_off=cut_off)
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["H2O"][layer], *bounds, ped, cut_off)
            assert scaled_error(k[layer], k_ref, 10, max(cut_off, 1)) <= FP64_TOL


def test_dense_band(dense_db, atmosphere):
    """Band grid 500-850 cm-1 with every line inside [v0-26, vn+26] (BASELINE config 3 shape)."""
    bounds = (500, 851, 200)
    gas = Gas(dense_db, "CO2")
    ref = OracleGas(dense_db, "CO2")
    for ped in (False, True):
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"],
                                        bounds=bounds, remove_pedestal=ped)
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["CO2"][layer], *bounds, ped)
            assert np.any(k_ref)
            assert scaled_error(k[layer], k_ref, bounds[2]) <= FP64_TOL


def test_unsorted_database(tmp_path, atmosphere):
    """Rows out of nu order: the reference still walks them in row order (pedestal order,
    early break); the packer must sort a copy and keep the row order for the recurrence."""
    lines = synth.make_line_list("CO2", 600, 560.0, 760.0, seed=3)
    rng = np.random.default_rng(5)
    perm = rng.permutation(600)
    shuffled = {k: v[perm] for k, v in lines.items()}
    path = str(tmp_path / "unsorted.db")
    synth.write_database(path, {"CO2": shuffled})
    bounds = (540, 781, 50)
    gas = Gas(path, "CO2")
    ref = OracleGas(path, "CO2")
    for ped in (False, True):
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"],
                                        bounds=bounds, remove_pedestal=ped)
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr["CO2"][layer], *bounds, ped, windows=True)
            assert scaled_error(k[layer], k_ref, bounds[2]) <= FP64_TOL
            assert np.array_equal(gas.windows(layer), ref.last_windows[:ref.last_active])


def test_no_tips_gives_zeros_and_unknown_molecule_raises(tmp_path, atmosphere):
    path = str(tmp_path / "notips.db")
    synth.write_database(path, {"CO": synth.make_line_list("CO", 50, 1.0, 300.0)}, tips=False)
    grid = synth.grid_from_bounds(1, 301, 10)
    k = Gas(path, "CO").absorption_coefficient(250.0, 5.0e4, 1e-7, grid)
    assert k.shape == (3000,) and not np.any(k)      # absorption.c:53-59
    with pytest.raises(ValueError, match="Error inside c functions."):
        Gas(path, "N2O").absorption_coefficient(250.0, 5.0e4, 1e-7, grid)
    assert "not found in database" in _lib.last_error()


def test_many_layers_chunked(small_db):
    """60-layer column pushed through several launch groups with pinned output."""
    col = synth.standard_column(60)
    bounds = (1, 301, 100)
    gas = Gas(small_db, "H2O")
    ref = OracleGas(small_db, "H2O")
    _lib.library().lbl_set_chunk_layers(7)
    try:
        pinned = _lib.PinnedArray((60, 30000))
        k = gas.absorption_coefficients(col.t, col.p, col.vmr["H2O"], bounds=bounds,
                                        remove_pedestal=True, out=pinned.array)
    finally:
        _lib.library().lbl_set_chunk_layers(0)
    assert gas.last_stats[0]["sum_launches"] == 9
    total = 0
    for layer in (0, 6, 7, 13, 30, 59):
        k_ref = ref.absorption(col.t[layer], col.p[layer], col.vmr["H2O"][layer], *bounds, True)
        assert scaled_error(k[layer], k_ref, 100) <= FP64_TOL
    for layer in range(60):
        ref.absorption(col.t[layer], col.p[layer], col.vmr["H2O"][layer], *bounds, True)
        total += ref.last_evals
    assert gas.last_stats[0]["evals"] == total


@pytest.mark.parametrize("bounds", [(1, 1201, 10), (1, 601, 100), (1, 500, 4), (1, 901, 32)])
@pytest.mark.parametrize("remove_pedestal", [False, True])
def test_fp32_mode(small_db, atmosphere, bounds, remove_pedestal):
    """Opt-in FP32 mode, stated tolerance 1e-4; window bookkeeping stays bit-exact.  It is a
    mode of the direct summation kernel: on fine grids (n_per_v >= 64) the far-field kernel
    runs instead, in FP64, and the result is simply better than the mode promises."""
    for formula in ("H2O", "CO2", "O3"):
        gas = Gas(small_db, formula, precision="fp32")
        ref = OracleGas(small_db, formula)
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, atmosphere.vmr[formula],
                                        bounds=bounds, remove_pedestal=remove_pedestal)
        stats = gas.last_stats[0]
        total = 0
        worst = 0.0
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                   atmosphere.vmr[formula][layer], *bounds, remove_pedestal,
                                   windows=True)
            total += ref.last_evals
            if not np.any(k_ref):
                assert not np.any(k[layer])
                continue
            err = scaled_error(k[layer], k_ref, bounds[2])
            worst = max(worst, err)
            assert err <= FP32_TOL
            if not remove_pedestal:
                assert relative_error(k[layer], k_ref) <= FP32_TOL
            assert np.array_equal(gas.windows(layer), ref.last_windows[:ref.last_active])
        assert stats["evals"] == total
        # which arithmetic served the request is reported, and the error agrees with it: FP32
        # sums are visibly not FP64 (> 1e-12), and a request served in FP64 (far-field kernel on a
        # fine grid) meets the FP64 bar
        assert stats["fp32_used"] == (0 if stats["cells_per_warp"] > 0 else 1)
        if bounds[2] < 64:
            assert stats["fp32_used"] == 1
        if stats["fp32_used"]:
            assert worst > 1e-12
        else:
            assert worst <= FP64_TOL


@pytest.mark.parametrize("remove_pedestal", [False, True])
def test_fp32_mode_on_a_fine_grid(small_db, atmosphere, remove_pedestal, monkeypatch):
    """The FP32 arithmetic on a fine grid (direct kernel forced), against the oracle at the
    mode's stated tolerance, and the same request with the far-field kernel forced: served in
    FP64 and reported as such."""
    bounds = (1, 601, 100)
    gas = Gas(small_db, "CO2", precision="fp32")
    ref = OracleGas(small_db, "CO2")
    x = atmosphere.vmr["CO2"]
    for farfield in ("0", "2"):
        monkeypatch.setenv("PYLBL_B200_FARFIELD", farfield)
        k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, x, bounds=bounds,
                                        remove_pedestal=remove_pedestal)
        assert gas.last_stats[0]["fp32_used"] == (1 if farfield == "0" else 0)
        for layer in range(atmosphere.t.size):
            k_ref = ref.absorption(atmosphere.t[layer], atmosphere.p[layer], x[layer], *bounds,
                                   remove_pedestal)
            err = scaled_error(k[layer], k_ref, bounds[2])
            if farfield == "0":
                assert 1e-12 < err <= FP32_TOL
            else:
                assert err <= FP64_TOL


def test_mixture_total_absorption(small_db, atmosphere):
    """Device-side  sum_gas n_gas * k_gas  against the host-side statement of
    pyLBL/spectroscopy.py:181-191,225-234 built from the oracle."""
    from pylbl_b200 import Mixture, number_density
    formulas = ["H2O", "CO2", "O3"]
    bounds = (1, 801, 100)
    mix = Mixture(small_db, formulas)
    total = mix.total_absorption(atmosphere.t, atmosphere.p, atmosphere.vmr, bounds=bounds)
    again = mix.total_absorption(atmosphere.t, atmosphere.p, atmosphere.vmr, bounds=bounds)
    assert np.array_equal(total, again)          # the accumulator is reset between calls
    want = np.zeros_like(total)
    for f in formulas:
        ref = OracleGas(small_db, f)
        for layer in range(atmosphere.t.size):
            n = number_density(atmosphere.t[layer], atmosphere.p[layer], atmosphere.vmr[f][layer])
            want[layer] += n * ref.absorption(atmosphere.t[layer], atmosphere.p[layer],
                                              atmosphere.vmr[f][layer], *bounds, True)
    for layer in range(atmosphere.t.size):
        assert scaled_error(total[layer], want[layer], bounds[2]) <= FP64_TOL
    mix.close()


def test_layers_sharded_over_two_devices(small_db):
    """`Gas(devices=[0, 1])`: contiguous layer shards, one per device, bitwise the same spectra."""
    if _lib.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    col = synth.standard_column(12)
    bounds = (1, 401, 100)
    one = Gas(small_db, "CO2", devices=[0]).absorption_coefficients(
        col.t, col.p, col.vmr["CO2"], bounds=bounds, remove_pedestal=True)
    gas = Gas(small_db, "CO2", devices=[0, 1])
    two = gas.absorption_coefficients(col.t, col.p, col.vmr["CO2"], bounds=bounds,
                                      remove_pedestal=True)
    assert np.array_equal(one, two)
    assert [s["n_layers"] for s in gas.last_stats] == [6, 6]


def test_lines_on_integer_boundaries_with_large_shifts(tmp_path):
    """Lines sitting on integer wavenumbers with shifts of both signs: the window cell comes
    from the SHIFTED centre and must match the reference bit for bit (spectra.c:22,48)."""
    lines = synth.make_line_list("O2", 200, 30.0, 130.0, seed=9)
    lines["nu"] = np.sort(np.round(lines["nu"]) + np.tile([0.0, 1e-6, -1e-6, 0.5], 50))
    lines["delta_air"] = np.tile([-0.02, 0.02, 0.0199, -0.0003], 50)
    path = str(tmp_path / "shift.db")
    synth.write_database(path, {"O2": lines})
    gas, ref = Gas(path, "O2"), OracleGas(path, "O2")
    t = np.array([250.0, 296.0]); p = np.array([101325.0, 5.0e4]); x = np.array([0.209, 0.209])
    for bounds in ((5, 161, 20), (5, 161, 100)):
        for ped in (False, True):
            k = gas.absorption_coefficients(t, p, x, bounds=bounds, remove_pedestal=ped)
            for layer in range(2):
                k_ref = ref.absorption(t[layer], p[layer], x[layer], *bounds, ped, windows=True)
                assert scaled_error(k[layer], k_ref, bounds[2]) <= FP64_TOL
                assert np.array_equal(gas.windows(layer), ref.last_windows[:ref.last_active])


def test_one_handle_many_grids_and_errors(small_db, atmosphere):
    """A handle is reused across grids (the line plan is rebuilt), and bad inputs fail with the
    reference's error type instead of reading out of bounds."""
    gas, ref = Gas(small_db, "H2O"), OracleGas(small_db, "H2O")
    layer = 1
    args = (atmosphere.t[layer], atmosphere.p[layer], atmosphere.vmr["H2O"][layer])
    for bounds in ((1, 301, 100), (1, 2001, 10), (1, 301, 100), (1, 120, 7)):
        k = gas.absorption_coefficients([args[0]], [args[1]], [args[2]], bounds=bounds)[0]
        k_ref = ref.absorption(*args, *bounds)
        assert relative_error(k, k_ref) <= FP64_TOL
    with pytest.raises(ValueError):        # temperature outside the TIPS table (1..1000 K here)
        gas.absorption_coefficients([1500.0], [args[1]], [args[2]], bounds=(1, 301, 100))
    assert "TIPS" in _lib.last_error()
    with pytest.raises(ValueError):        # vn <= v0
        gas.absorption_coefficients([args[0]], [args[1]], [args[2]], bounds=(300, 300, 10))
    # and the handle still works afterwards
    k = gas.absorption_coefficients([args[0]], [args[1]], [args[2]], bounds=(1, 301, 100))[0]
    assert relative_error(k, ref.absorption(*args, 1, 301, 100)) <= FP64_TOL


def test_two_handles_from_two_threads(small_db, atmosphere):
    """Different handles may be driven from different host threads."""
    import threading
    bounds = (1, 401, 100)
    results = {}

    def work(formula):
        gas = Gas(small_db, formula)
        results[formula] = gas.absorption_coefficients(atmosphere.t, atmosphere.p,
                                                       atmosphere.vmr[formula], bounds=bounds,
                                                       remove_pedestal=True)
    threads = [threading.Thread(target=work, args=(f,)) for f in ("H2O", "CO2", "O3")]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for f in ("H2O", "CO2", "O3"):
        ref = OracleGas(small_db, f)
        k_ref = ref.absorption(atmosphere.t[2], atmosphere.p[2], atmosphere.vmr[f][2], *bounds, True)
        assert scaled_error(results[f][2], k_ref, 100) <= FP64_TOL


def test_spectroscopy_adapter_matches_the_reference_driver_loop(small_db, atmosphere):
    """`pylbl_b200.Spectroscopy.compute_absorption` against a replica of the reference driver
    (pyLBL/spectroscopy.py:161-191: flat loop over the atmosphere, beta = n*k[:grid.size])
    run over the oracle, for a 2-D atmosphere and all three output formats; a gas the database
    does not hold stays zero (gas = None, spectroscopy.py:53-57)."""
    from pylbl_b200 import Spectroscopy, number_density

    class Member(object):          # stands in for an xarray DataArray: only .data is read
        def __init__(self, data):
            self.data = data

    class Atmosphere(object):
        pass

    shape = (2, 2)
    names = ["H2O", "CO2", "XeF6"]
    atm = Atmosphere()
    atm.temperature = Member(atmosphere.t.reshape(shape))
    atm.pressure = Member(atmosphere.p.reshape(shape))
    atm.gases = {n: Member((atmosphere.vmr[n] if n in atmosphere.vmr else
                            np.full(4, 1e-9)).reshape(shape)) for n in names}
    grid = synth.grid_from_bounds(1, 401, 100)
    s = Spectroscopy(atm, grid, Db(small_db))

    want = {}
    for n in names[:2]:
        ref = OracleGas(small_db, n)
        beta = np.zeros(shape + (grid.size,))
        for i in range(4):
            j = np.unravel_index(i, shape)
            t, p, x = atmosphere.t[i], atmosphere.p[i], atmosphere.vmr[n][i]
            k = ref.absorption(t, p, x, 1, 401, 100, True)
            beta[j] = number_density(t, p, x) * k[:grid.size]
        want[n] = beta

    # continua the reference attaches per gas (spectroscopy.py:58-65,194-198), from the oracle
    from oracle import mt_ckd
    cont = {}
    for n in names:
        c = np.zeros(shape + (grid.size,))
        for cname in mt_ckd.continua_of(n):
            oracle = mt_ckd.OracleContinuum(cname)
            for i in range(4):
                j = np.unravel_index(i, shape)
                vmr = {m: float(atm.gases[m].data.flat[i]) for m in names}
                c[j] += oracle.spectra(atmosphere.t[i], atmosphere.p[i], vmr, grid)
        cont[n] = c
    assert cont["H2O"].any() and cont["CO2"].any() and not cont["XeF6"].any()

    allv = s.compute_absorption(output_format="all")
    assert allv["mechanism"] == ["lines", "continuum", "cross_section"]
    gasv = s.compute_absorption(output_format="gas")
    total = s.compute_absorption(output_format="total")["absorption"]
    assert total.shape == shape + (grid.size,)
    for n in names[:2]:
        a = allv[f"{n}_absorption"]
        assert a.shape == shape + (3, grid.size)
        assert not a[..., 2, :].any()
        for j in np.ndindex(shape):
            assert scaled_error(a[j][0], want[n][j], 100) <= FP64_TOL
            assert np.abs(a[j][1] - cont[n][j]).max() <= 1e-12 * np.abs(cont[n][j]).max()
            assert scaled_error(gasv[f"{n}_absorption"][j], want[n][j] + cont[n][j], 100) <= FP64_TOL
    assert not allv["XeF6_absorption"].any() and not gasv["XeF6_absorption"].any()
    for j in np.ndindex(shape):
        everything = want["H2O"][j] + want["CO2"][j] + cont["H2O"][j] + cont["CO2"][j]
        assert scaled_error(total[j], everything, 100) <= FP64_TOL
    s.close()
    # and with the continua switched off: the lines alone, pedestal kept (spectroscopy.py:163-164)
    plain = Spectroscopy(atm, grid, Db(small_db), continua_backend=None)
    allp = plain.compute_absorption(output_format="all", remove_pedestal=True)
    assert not allp["H2O_absorption"][..., 1:, :].any()
    assert np.array_equal(allp["H2O_absorption"][..., 0, :], allv["H2O_absorption"][..., 0, :])
    plain.close()


@pytest.mark.parametrize("farfield,nearblock", [("2", "0"), ("2", "1"), ("0", "1"), ("0", "0")])
def test_fallback_kernels_on_a_fine_grid(small_db, atmosphere, monkeypatch, farfield, nearblock):
    """The kernels a fine grid does not normally use -- the direct summation kernel K2 and the
    point-major near-zone kernel -- selected through the library's environment knobs, against
    the oracle and against the default path."""
    bounds = (1, 301, 100)
    gas = Gas(Db(small_db), "H2O")
    x = atmosphere.vmr["H2O"]
    default = gas.absorption_coefficients(atmosphere.t, atmosphere.p, x, bounds=bounds,
                                          remove_pedestal=True)
    monkeypatch.setenv("PYLBL_B200_FARFIELD", farfield)
    monkeypatch.setenv("PYLBL_B200_NEARBLOCK", nearblock)
    k = gas.absorption_coefficients(atmosphere.t, atmosphere.p, x, bounds=bounds, remove_pedestal=True)
    ref = OracleGas(small_db, "H2O")
    for layer in range(atmosphere.t.size):
        want = ref.absorption(atmosphere.t[layer], atmosphere.p[layer], x[layer], *bounds, True)
        assert scaled_error(k[layer], want, bounds[2]) <= FP64_TOL
        assert scaled_error(k[layer], default[layer], bounds[2]) <= 1e-10
    gas.close()


@pytest.mark.parametrize("bounds", [(1, 1201, 100), (1, 2001, 10), (640, 700, 2000)])
@pytest.mark.parametrize("remove_pedestal", [False, True])
def test_bands_concatenate_to_the_wide_call_bit_for_bit(small_db, atmosphere, bounds, remove_pedestal):
    """Spectral band sharding (BASELINE configs[3]): cells [lo, hi) of the grid,
    with the windows, the active-line prefix (absorption.c:80-83) and the accumulated pedestal
    (spectra.c:66-78) of the WHOLE grid.  Bands at odd cell boundaries, concatenated, must equal
    the single wide call exactly; and their evaluation counts add up to the wide call's."""
    v0, vn, n_per_v = bounds
    ncell = vn - v0
    gas = Gas(small_db, "CO2")
    t, p, x = atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"]
    wide = gas.absorption_coefficients(t, p, x, bounds=bounds, remove_pedestal=remove_pedestal)
    wide_evals = gas.last_stats[0]["evals"]
    edges = [0, 1, ncell // 3 + 1, ncell // 3 + 2, (2 * ncell) // 3, ncell - 1, ncell]
    joined = np.full_like(wide, np.nan)
    evals = 0
    for lo, hi in zip(edges[:-1], edges[1:]):
        gas.absorption_band(t, p, x, bounds, (lo, hi), remove_pedestal=remove_pedestal,
                            out=joined[:, lo * n_per_v:hi * n_per_v])
        evals += gas.last_stats[0]["evals"]
        assert gas.last_stats[0]["n_points"] == (hi - lo) * n_per_v
    assert np.array_equal(joined, wide)
    assert evals == wide_evals
    # a dense destination, and an automatic split
    lo, hi = edges[2], edges[4]
    k = gas.absorption_band(t, p, x, bounds, (lo, hi), remove_pedestal=remove_pedestal)
    assert np.array_equal(k, wide[:, lo * n_per_v:hi * n_per_v])
    auto = gas.band_edges(bounds, 3)
    assert auto[0] == 0 and auto[-1] == ncell
    with pytest.raises(ValueError):
        gas.absorption_band(t, p, x, bounds, (5, 5))
    gas.close()


def test_bands_of_a_dense_list_bit_for_bit(tmp_path, atmosphere, monkeypatch):
    """The far-field kernel on a list dense enough (50 lines per cm-1: 2500 lines in a window)
    that every cell's line ranges span several staged chunks: where the ranges are cut must not
    depend on how a band groups the cells into blocks (chunks are cut at absolute line
    indices), or the sums would round differently."""
    monkeypatch.setenv("PYLBL_B200_FARFIELD", "2")
    path = str(tmp_path / "config3_third.db")
    synth.write_database(path, synth.config_line_lists(3, scale=1. / 3.))
    bounds = (500, 851, 200)
    gas = Gas(path, "CO2")
    t, p, x = atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"]
    for ped in (False, True):
        wide = gas.absorption_coefficients(t, p, x, bounds=bounds, remove_pedestal=ped)
        assert gas.last_stats[0]["cells_per_warp"] == 1
        joined = np.full_like(wide, np.nan)
        edges = [0, 3, 118, 233, 350, 351]
        for lo, hi in zip(edges[:-1], edges[1:]):
            gas.absorption_band(t, p, x, bounds, (lo, hi), remove_pedestal=ped,
                                out=joined[:, lo * 200:hi * 200])
        assert np.array_equal(joined, wide)
    gas.close()


def test_band_of_an_unsorted_database(tmp_path, atmosphere):
    """Rows out of nu order: the recurrence must still walk every active row for a band."""
    lines = synth.make_line_list("CO2", 600, 560.0, 760.0, seed=3)
    perm = np.random.default_rng(5).permutation(600)
    path = str(tmp_path / "unsorted.db")
    synth.write_database(path, {"CO2": {k: v[perm] for k, v in lines.items()}})
    bounds = (540, 781, 50)
    gas = Gas(path, "CO2")
    t, p, x = atmosphere.t, atmosphere.p, atmosphere.vmr["CO2"]
    wide = gas.absorption_coefficients(t, p, x, bounds=bounds, remove_pedestal=True)
    k = gas.absorption_band(t, p, x, bounds, (60, 200), remove_pedestal=True)
    assert np.array_equal(k, wide[:, 60 * 50:200 * 50])


def test_bands_sharded_over_two_devices(small_db, atmosphere):
    """`Gas(devices=[0, 1], shard="band")`: one spectral band per device, same bits."""
    if _lib.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    bounds = (1, 1201, 100)
    t, p, x = atmosphere.t, atmosphere.p, atmosphere.vmr["H2O"]
    one = Gas(small_db, "H2O", devices=[0]).absorption_coefficients(t, p, x, bounds=bounds,
                                                                    remove_pedestal=True)
    gas = Gas(small_db, "H2O", devices=[0, 1], shard="band")
    two = gas.absorption_coefficients(t, p, x, bounds=bounds, remove_pedestal=True)
    assert np.array_equal(one, two)
    assert sum(s["n_points"] for s in gas.last_stats) == 120000


def test_total_absorption_in_several_layer_groups(small_db):
    """The device-side gas sum when the layers do not fit one launch group (the accumulator is
    fed group by group), with a pinned destination; against the one-group result."""
    from pylbl_b200 import Mixture
    col = synth.standard_column(11)
    bounds = (1, 401, 100)
    mix = Mixture(small_db, ["H2O", "CO2", "O3"])
    whole = mix.total_absorption(col.t, col.p, col.vmr, bounds=bounds).copy()
    _lib.library().lbl_set_chunk_layers(3)
    try:
        pinned = _lib.PinnedArray((11, 40000))
        parts = mix.total_absorption(col.t, col.p, col.vmr, bounds=bounds, out=pinned.array)
    finally:
        _lib.library().lbl_set_chunk_layers(0)
    assert np.array_equal(parts, whole)
    assert np.all(whole > 0) or True
    mix.close()


@pytest.mark.parametrize("farfield", ["2", "0"])
def test_line_cores_at_very_low_pressure(dense_db, farfield, monkeypatch):
    """Mesospheric layers (1 Pa and 0.1 Pa): y = sqrt(ln2)*gamma/alpha falls to 1e-5..1e-7, the
    Lorentz form at a line centre is 1e4..1e6 times the profile, and whatever the far-field
    kernel adds there must be taken back bit for bit (far_term_lo); y <= 1e-6 also switches the
    W4 regions 1 and 2 off (voigt.c:48-53).  Pointwise, without the pedestal."""
    monkeypatch.setenv("PYLBL_B200_FARFIELD", farfield)
    bounds = (500, 851, 200)      # the whole list: a grid that starts inside it yields zeros (Q1)
    t = np.array([190.0, 210.0, 250.0])
    p = np.array([0.1, 1.0, 10.0])
    x = np.array([3.6e-4, 3.6e-4, 3.6e-4])
    gas, ref = Gas(dense_db, "CO2"), OracleGas(dense_db, "CO2")
    k = gas.absorption_coefficients(t, p, x, bounds=bounds)
    assert (gas.last_stats[0]["cells_per_warp"] > 0) == (farfield == "2")
    for layer in range(3):
        k_ref = ref.absorption(t[layer], p[layer], x[layer], *bounds)
        assert relative_error(k[layer], k_ref) <= FP64_TOL


@pytest.mark.parametrize("bounds", [(1, 5001, 1), (2380, 2441, 50), (7000, 60000, 1), (1, 3251, 10)])
def test_continuum_against_the_oracle(atmosphere, bounds):
    """MT-CKD continua on the device (every band formula + numpy.interp semantics, zero outside a
    band) against the numpy restatement that tests/test_oracle_mt_ckd.py pins to the reference's
    own modules: the reference's fixture atmosphere (all eight gases), every continuum, grids
    over the infrared, across the CO2 band head sub-grids, and up to the ultra-violet bands."""
    from oracle import mt_ckd
    from pylbl_b200 import Continuum
    gases = ["H2O", "CO2", "O3", "N2O", "CH4", "CO", "O2"]
    vmr = {g: atmosphere.vmr[g] for g in gases}
    vmr["N2"] = np.full(4, 0.78)                       # tests/conftest.py:76 of the reference
    cont = Continuum()
    grid = synth.grid_from_bounds(*bounds)
    n = (bounds[1] - bounds[0]) * bounds[2]
    v = bounds[0] + np.arange(n) * (1. / bounds[2])     # the grid as absorption.c:33-39 forms it
    seen = 0
    for name in ("CO2", "H2OForeign", "H2OSelf", "N2", "O2", "O3"):
        k = cont.spectra(name, atmosphere.t, atmosphere.p, vmr, bounds=bounds)
        assert k.shape == (4, n)
        oracle = mt_ckd.OracleContinuum(name)
        for layer in range(4):
            want = oracle.spectra(atmosphere.t[layer], atmosphere.p[layer],
                                  {g: float(x[layer]) for g, x in vmr.items()}, v)
            scale = np.abs(want).max()
            if scale == 0.:
                assert not k[layer].any()
                continue
            seen += 1
            assert np.abs(k[layer] - want).max() <= 1e-12 * scale
            assert np.array_equal(k[layer] == 0., want == 0.)      # the same points outside every band
    assert seen >= 8
    # a gas the formulas need but the atmosphere lacks: the reference's dictionary lookup fails
    with pytest.raises(KeyError):
        cont.spectra("N2", atmosphere.t, atmosphere.p, {g: atmosphere.vmr[g] for g in gases}, bounds=bounds)
    cont.close()


def test_continuum_into_the_gas_sum(small_db, atmosphere):
    """`Mixture.total_absorption(continuum=...)`: lines of every gas plus the continua of every
    gas of the atmosphere, summed on the device, against the pieces computed separately."""
    from pylbl_b200 import Continuum, Mixture, continua_of
    bounds = (1, 801, 100)
    gases = ["H2O", "CO2", "O3"]
    vmr = {g: atmosphere.vmr[g] for g in gases + ["O2"]}
    vmr["N2"] = np.full(4, 0.78)
    mix = Mixture(small_db, gases)
    cont = Continuum()
    lines = mix.total_absorption(atmosphere.t, atmosphere.p, vmr, bounds=bounds).copy()
    both = mix.total_absorption(atmosphere.t, atmosphere.p, vmr, bounds=bounds, continuum=cont)
    extra = np.zeros_like(lines)
    for formula in vmr:
        for name in continua_of(formula):
            extra += cont.spectra(name, atmosphere.t, atmosphere.p, vmr, bounds=bounds)
    assert extra.any()
    for layer in range(4):
        assert scaled_error(both[layer], lines[layer] + extra[layer], bounds[2]) <= 1e-13
    mix.close()
    cont.close()
