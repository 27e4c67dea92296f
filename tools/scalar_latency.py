"""Latency of the scalar plugin call (what an unmodified pyLBL driver makes per (gas, layer)):
python tools/scalar_latency.py"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pylbl_b200 import Gas, synth

db = bench.database_path(0, lambda: None)
v0, vn, npv = synth.config_grid(2)
grid = synth.grid_from_bounds(v0, vn, npv)
col = synth.standard_column(60)
for f in ("CO", "N2O", "CO2"):
    g = Gas(db, f, devices=[0])
    for ped in (False, True):
        for layer in (0, 1):
            g.absorption_coefficient(col.t[layer], col.p[layer], col.vmr[f][layer], grid, remove_pedestal=ped)
        t0 = time.perf_counter()
        n = 10
        for layer in range(n):
            g.absorption_coefficient(col.t[layer], col.p[layer], col.vmr[f][layer], grid, remove_pedestal=ped)
        dt = (time.perf_counter() - t0) / n
        s = g.last_stats[0]
        print(f"{f:4s} pedestal={ped!s:5s} {dt*1e3:7.2f} ms per call  (device {s['total_ms']:.2f} ms: "
              f"scale {s['scale_ms']:.2f} sum {s['sum_ms']:.2f} near {s['fixup_ms']:.2f} pedestal {s['pedestal_ms']:.2f})")
    g.close()
