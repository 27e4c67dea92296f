"""Cost of equal-width spectral bands of BASELINE configs[3] (1M lines, 10-3500 cm-1 @0.001), to
calibrate the cost model behind lbl_gas_band_edges: python tools/band_cost.py [bands]"""
import ctypes
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from pylbl_b200 import Gas, _lib, synth

n_bands = int(sys.argv[1]) if len(sys.argv) > 1 else 16
db = bench.aux_database(4, 0, lambda: None)
bounds = synth.config_grid(4)
v0, vn, npv = bounds
col = synth.standard_column(60)
gas = Gas(db, "XX", devices=[0])
lines = synth.config_line_lists(4)["XX"]["nu"]
edges = np.linspace(0, vn - v0, n_bands + 1).astype(int)
auto = gas.band_edges(bounds, 8)
print(json.dumps({"automatic_edges_8": [int(x) for x in auto]}))
for b in range(n_bands):
    lo, hi = int(edges[b]), int(edges[b + 1])
    for rep in range(2):
        gas.absorption_band(col.t, col.p, col.vmr["XX"], bounds, (lo, hi), remove_pedestal=True,
                            out=None if rep else None)
    s = gas.last_stats[0]
    inside = int(np.sum((lines >= v0 + lo) & (lines < v0 + hi)))
    print(json.dumps({"band": [lo, hi], "lines_in_band": inside, "centre": v0 + 0.5 * (lo + hi),
                      "sum_ms": round(s["sum_ms"], 3), "near_ms": round(s["fixup_ms"], 3),
                      "scale_ms": round(s["scale_ms"], 3), "pedestal_ms": round(s["pedestal_ms"], 3),
                      "total_ms": round(s["total_ms"], 3)}))
for b in range(8):
    lo, hi = int(auto[b]), int(auto[b + 1])
    for rep in range(2):
        gas.absorption_band(col.t, col.p, col.vmr["XX"], bounds, (lo, hi), remove_pedestal=True)
    s = gas.last_stats[0]
    print(json.dumps({"auto_band": [lo, hi], "sum_ms": round(s["sum_ms"], 3), "near_ms": round(s["fixup_ms"], 3),
                      "scale_ms": round(s["scale_ms"], 3), "pedestal_ms": round(s["pedestal_ms"], 3),
                      "total_ms": round(s["total_ms"], 3)}))
