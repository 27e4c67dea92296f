"""CPU restatement of pyLBL's MT-CKD continuum plugin.  TEST INFRASTRUCTURE ONLY (see __init__).

Follows, function by function, /root/reference/pyLBL/mt_ckd/: utils.py (number densities :18-44,
radiation term :47-61, sub-grid placement :64-81, interpolation onto the caller's grid :157-174)
and the band formulas of carbon_dioxide.py, water_vapor.py, nitrogen.py, oxygen.py, ozone.py.
The coefficients come from pylbl_b200/data/mt_ckd.npz (tools/convert_mt_ckd.py).  Pinned against
tests/golden/mt_ckd_reference.npz, which tests/golden/make_mt_ckd_golden.py produced by running
the reference's own modules (tests/test_oracle_mt_ckd.py).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

TABLE = Path(__file__).resolve().parent.parent / "pylbl_b200" / "data" / "mt_ckd.npz"

LOSCHMIDT = 2.6867775e19   # utils.py:7
P0 = 1013.25               # utils.py:8
C2 = 1.4387752             # utils.py:9
T0 = 296.                  # utils.py:10
T273 = 273.15              # utils.py:11
M_TO_CM = 100.             # utils.py:12
PA_TO_MB = 0.01            # utils.py:13


def dry_air_number_density(p, t, vmr):                  # utils.py:33-44
    return LOSCHMIDT * (p / P0) * (T273 / t) * (1. - vmr["H2O"])


def air_number_density(p, t, vmr):                      # utils.py:18-30
    return sum([dry_air_number_density(p, t, vmr) * x for x in vmr.values()])


def radiation_term(w, t):                               # utils.py:47-61
    x = w / (t / C2)
    r = np.where(x <= 0.01, 0.5 * x * w, w)
    return np.where(x <= 10., w * (1. - np.exp(-x)) / (1. + np.exp(-x)), r)


class Table(object):
    def __init__(self, path=TABLE):
        z = np.load(path)
        self.data = {k: z[k] for k in z.files if not k.endswith("__grid")}
        self.grid = {k[:-6]: z[k] for k in z.files if k.endswith("__grid")}

    def wavenumbers(self, name):                        # utils.py:138-144
        lower, _, res = self.grid[name]
        return np.asarray([lower + i * res for i in range(self.data[name].size)])

    def placed(self, host, name, fill):
        """`name` laid over the grid of `host`, `fill` elsewhere (utils.py:64-81 and its callers)."""
        g, s = self.grid[host], self.grid[name]
        assert g[2] == s[2] and g[0] <= s[0] and g[1] >= s[1]
        lower = int((s[0] - g[0]) / g[2])
        upper = int((s[1] - g[0]) / g[2])
        out = np.full(self.data[host].size, fill)
        out[lower:upper + 1] = self.data[name]
        return out, lower, upper


def bands(table):
    """{continuum name: [(band grid, function(t, p_mb, vmr) -> band spectrum), ...]} in the order
    of the reference's ``bands`` lists."""
    d, w = table.data, table.wavenumbers

    # carbon_dioxide.py:27-45
    w_co2 = w("bfco2")
    tcorr, _, _ = table.placed("bfco2", "tdep_bandhead", 1.)
    xfac, _, _ = table.placed("bfco2", "x_factor_co2", 1.)

    def co2(t, p, v):
        n = dry_air_number_density(p, t, v) * v["CO2"]
        return n * 1.e-20 * (p / P0) * (T0 / t) * radiation_term(w_co2, t) * xfac * \
            np.power(t / 246., tcorr) * d["bfco2"]

    # water_vapor.py:12-35
    w_self = w("bs296")

    def h2o_self(t, p, v):
        tf = (t - T0) / (260. - T0)
        nh2o = dry_air_number_density(p, t, v) * v["H2O"]
        n = air_number_density(p, t, v)
        return nh2o * (nh2o / n) * (p / P0) * (T0 / t) * 1.e-20 * radiation_term(w_self, t) * \
            d["bs296"] * np.power(d["bs260"] / d["bs296"], tf)

    # water_vapor.py:43-79
    w_for = w("bfh2o")
    x, lower, upper = table.placed("bfh2o", "xfac_rhu", 0.)
    scale = np.zeros(d["bfh2o"].size)
    scale[lower + 1:upper + 1] = d["xfac_rhu"][1:]
    scale[lower] = scale[lower + 1]
    u = upper + 1
    ww = w_for[u:]
    vdelsq1 = (ww - 255.67) * (ww - 255.67)
    vf1 = np.power((ww - 255.67) / 57.83, 8)
    vdelmsq1 = (ww + 255.67) * (ww + 255.67)
    vmf1 = np.power((ww + 255.67) / 57.83, 8)
    vf2 = np.power(ww / 630., 8)
    scale[u:] = 1. + (0.06 - 0.42 * ((57600. / (vdelsq1 + 57600. + vf1)) +
                                     (57600. / (vdelmsq1 + 57600. + vmf1)))) / (1. + 0.3 * vf2)

    def h2o_foreign(t, p, v):
        nh2o = dry_air_number_density(p, t, v) * v["H2O"]
        n = air_number_density(p, t, v)
        return (1. - (nh2o / n)) * (p / P0) * (T0 / t) * 1.e-20 * nh2o * radiation_term(w_for, t) * \
            scale * d["bfh2o"]

    # nitrogen.py:15-38
    w_rot = w("ct_296")

    def n2_rot(t, p, v):
        nn2 = dry_air_number_density(p, t, v) * v["N2"]
        tau = (nn2 / LOSCHMIDT) * (p / P0) * (T273 / t)
        f = (t - T0) / (220. - T0)
        c = d["ct_296"] * np.power(d["ct_220"] / d["ct_296"], f)
        s = d["sf_296"] * np.power(d["sf_220"] / d["sf_296"], f)
        fo2 = (s - 1.) * v["N2"] / v["O2"]
        return tau * radiation_term(w_rot, t) * c * (v["N2"] + fo2 * v["O2"] + v["H2O"])

    # nitrogen.py:41-63
    w_fun = w("xn2_272")

    def n2_fund(t, p, v):
        nn2 = dry_air_number_density(p, t, v) * v["N2"]
        tau = (nn2 / LOSCHMIDT) * (p / P0) * (T273 / t)
        xt = (1. / t - 1. / 272.) / (1. / 228. - 1. / 272.)
        ao2 = 1.294 - 0.4545 * t / T0
        c0 = np.zeros(d["xn2_272"].size)
        with np.errstate(divide="ignore", invalid="ignore"):
            c0[1:-1] = d["xn2_272"][1:-1] * np.power(d["xn2_228"][1:-1] / d["xn2_272"][1:-1], xt)
        c0 = c0 / w_fun
        c1 = ao2 * c0
        c2 = (9. / 7.) * d["a_h2o"] * c0
        return tau * radiation_term(w_fun, t) * (c0 * v["N2"] + v["O2"] * c1 + v["H2O"] * c2)

    # nitrogen.py:66-79
    w_ovt = w("xn2")

    def n2_overtone(t, p, v):
        nn2 = dry_air_number_density(p, t, v) * v["N2"]
        tau = (nn2 / LOSCHMIDT) * (p / P0) * (T273 / t) * (v["N2"] + v["O2"] + v["H2O"])
        return tau * radiation_term(w_ovt, t) * d["xn2"] / w_ovt

    # oxygen.py:19-35
    w_o2f = w("o2_f")

    def o2_fund(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        tau = no2 * 1.e-20 * (p / P0) * (T273 / t)
        xkt = (1. / T0) - (1. / t)
        factor = (1.e20 / LOSCHMIDT)
        return tau * radiation_term(w_o2f, t) * factor * d["o2_f"] * np.exp(d["o2_t"] * xkt) / w_o2f

    # oxygen.py:38-54
    w_nir = w("o2_inf1")

    def o2_nir(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        ao2, an2 = 1. / 0.446, 0.3 / 0.446
        tau = (no2 / LOSCHMIDT) * (p / P0) * (T273 / t) * (ao2 * v["O2"] + an2 * v["N2"] + v["H2O"])
        return tau * radiation_term(w_nir, t) * d["o2_inf1"] / w_nir

    # oxygen.py:57-81
    w_nir2 = np.arange(9100., 11002., 2.)
    nir2 = np.zeros(w_nir2.size)
    hw1, hw2 = 58.96, 45.04
    for i in range(w_nir2.size):
        dv1 = w_nir2[i] - 9375.
        dv2 = w_nir2[i] - 9439.
        damp1 = np.exp(dv1 / 176.1) if dv1 < 0. else 1.
        damp2 = np.exp(dv2 / 176.1) if dv2 < 0. else 1.
        o2inf = 0.31831 * (((1.166e-04 * damp1 / hw1) / (1. + (dv1 / hw1) * (dv1 / hw1))) +
                           ((3.086e-05 * damp2 / hw2) / (1. + (dv2 / hw2) * (dv2 / hw2)))) * 1.054
        nir2[i] = o2inf / w_nir2[i]

    def o2_nir2(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        n = air_number_density(p, t, v)
        adj = (no2 / n) * (1. / v["O2"]) * no2 * 1.e-20 * (p / P0) * (T0 / t)
        return adj * radiation_term(w_nir2, t) * nir2

    # oxygen.py:84-95
    w_nir3 = w("o2_inf3")

    def o2_nir3(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        tau = (no2 / LOSCHMIDT) * (p / P0) * (T273 / t)
        return tau * radiation_term(w_nir3, t) * d["o2_inf3"] / w_nir3

    # oxygen.py:98-111
    w_vis = w("o2_invis")

    def o2_vis(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        n = air_number_density(p, t, v)
        adj = (no2 / n) * no2 * 1.e-20 * (p / P0) * (T273 / t)
        factor = 1. / (LOSCHMIDT * 1.e-20 * (55. * T273 / T0) * (55. * T273 / T0) * 89.5)
        return adj * radiation_term(w_vis, t) * factor * d["o2_invis"] / w_vis

    # oxygen.py:114-135
    w_hz = np.arange(36000., 100010., 10.)
    hz = np.zeros(w_hz.size)
    for i in range(w_hz.size):
        if w_hz[i] <= 36000.:
            hz[i] = 0.
        else:
            corr = ((40000. - w_hz[i]) / 4000.) * 7.917e-7 if w_hz[i] <= 40000. else 0.
            y = w_hz[i] / 48811.0
            hz[i] = 6.884e-4 * y * np.exp(-69.738 * np.power(np.log(y), 2)) - corr

    def o2_herzberg(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        factor = 1. + 0.83 * (p / P0) * (T273 / t)
        return 1.e-20 * no2 * radiation_term(w_hz, t) * factor * hz / w_hz

    # oxygen.py:138-148
    w_uv = w("o2_infuv")

    def o2_uv(t, p, v):
        no2 = dry_air_number_density(p, t, v) * v["O2"]
        return 1.e-20 * no2 * radiation_term(w_uv, t) * d["o2_infuv"] / w_uv

    # ozone.py:12-30
    w_cw = w("x_o3")

    def o3_cw(t, p, v):
        no3 = dry_air_number_density(p, t, v) * v["O3"]
        dt = t - T273
        return 1.e-20 * no3 * radiation_term(w_cw, t) * (d["x_o3"] + d["y_o3"] * dt + d["z_o3"] * dt * dt) / w_cw

    # ozone.py:33-53
    w_hh = w("o3_hh0")

    def o3_hh(t, p, v):
        no3 = dry_air_number_density(p, t, v) * v["O3"]
        dt = t - T273
        return 1.e-20 * no3 * radiation_term(w_hh, t) * (d["o3_hh0"] / w_hh) * \
            (1. + d["o3_hh1"] * dt + d["o3_hh2"] * dt * dt)

    # ozone.py:56-70
    w_o3uv = w("o3_huv")

    def o3_uv(t, p, v):
        no3 = dry_air_number_density(p, t, v) * v["O3"]
        return no3 * radiation_term(w_o3uv, t) * d["o3_huv"] / w_o3uv

    return {
        "CO2": [(w_co2, co2)],
        "H2OForeign": [(w_for, h2o_foreign)],
        "H2OSelf": [(w_self, h2o_self)],
        "N2": [(w_rot, n2_rot), (w_fun, n2_fund), (w_ovt, n2_overtone)],
        "O2": [(w_o2f, o2_fund), (w_nir, o2_nir), (w_nir2, o2_nir2), (w_nir3, o2_nir3),
               (w_vis, o2_vis), (w_hz, o2_herzberg), (w_uv, o2_uv)],
        "O3": [(w_cw, o3_cw), (w_hh, o3_hh), (w_o3uv, o3_uv)],
    }


class OracleContinuum(object):
    """``BandedContinuum`` (utils.py:147-174) for one continuum name."""
    _tables = {}

    def __init__(self, name, path=TABLE):
        if path not in self._tables:
            self._tables[path] = bands(Table(path))
        self.bands = self._tables[path][name]

    def spectra(self, temperature, pressure, vmr, grid):
        """Continuum extinction [m-1] on `grid`; pressure in Pa (utils.py:157-174)."""
        s = np.zeros(grid.size)
        for w, band in self.bands:
            s += np.interp(grid, w, band(temperature, pressure * PA_TO_MB, vmr), left=0., right=0.) * M_TO_CM
        return s


def continua_of(formula):
    """Continuum names the driver attaches to a gas (spectroscopy.py:58-65)."""
    if formula == "H2O":
        return ["H2OForeign", "H2OSelf"]
    return [formula] if formula in ("CO2", "N2", "O2", "O3") else []
