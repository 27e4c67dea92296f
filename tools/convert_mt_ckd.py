"""One-off converter: the reference's MT-CKD coefficient file (pyLBL/mt_ckd/mt-ckd.nc, netCDF-4 /
HDF5) -> pylbl_b200/data/mt_ckd.npz, a flat table the backend can load without netCDF4 or h5py
(neither exists in the build image; tools/hdf5_min.py reads the file).

    python tools/convert_mt_ckd.py [/root/reference/pyLBL/mt_ckd/mt-ckd.nc]

For every variable the band modules read (pyLBL/mt_ckd/*.py via utils.Spectrum, utils.py:117-144)
the table holds `<name>` (float64 data) and `<name>__grid` = (wavenumber_lower_bound,
wavenumber_upper_bound, wavenumber_resolution).  Values are copied bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import hdf5_min  # noqa: E402

VARIABLES = ["bfco2", "tdep_bandhead", "x_factor_co2",                       # carbon_dioxide.py
             "bs296", "bs260", "bfh2o", "xfac_rhu",                          # water_vapor.py
             "ct_296", "sf_296", "ct_220", "sf_220", "xn2_272", "xn2_228", "a_h2o", "xn2",  # nitrogen.py
             "o2_f", "o2_t", "o2_inf1", "o2_inf3", "o2_invis", "o2_infuv",   # oxygen.py
             "x_o3", "y_o3", "z_o3", "o3_hh0", "o3_hh1", "o3_hh2", "o3_huv"]  # ozone.py


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pyLBL/mt_ckd/mt-ckd.nc"
    out = Path(__file__).resolve().parent.parent / "pylbl_b200" / "data" / "mt_ckd.npz"
    datasets = hdf5_min.read_file(src)
    table = {}
    for name in VARIABLES:
        d = datasets[name]
        if d.data is None or d.data.dtype != np.float64 or d.data.ndim != 1:
            raise SystemExit(f"{name}: unexpected dataset {d.dims} {d.dtype}")
        grid = [float(np.asarray(d.attrs[f"wavenumber_{x}"]).ravel()[0])
                for x in ("lower_bound", "upper_bound", "resolution")]
        table[name] = d.data
        table[name + "__grid"] = np.asarray(grid)
    np.savez_compressed(out, **table)
    print(out, out.stat().st_size, "bytes,", len(VARIABLES), "variables")


if __name__ == "__main__":
    main()
