"""Per-gas kernel timing of one bench step (CUDA-event times from lbl_stats)."""
import sys, json, ctypes, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from pylbl_b200 import Gas, synth, _lib

ped = "--no-pedestal" not in sys.argv
db = bench.database_path(0, lambda: None)
config = 5 if "--config5" in sys.argv else 2       # same line lists; 5: the 0.1 cm-1 grid
bounds = synth.config_grid(config)
col = synth.standard_column(60)
rows = []
for f in bench.GASES:
    g = Gas(db, f, devices=[0])
    for rep in range(3):
        g.absorption_coefficients(col.t, col.p, col.vmr[f], bounds=bounds, remove_pedestal=ped, to_host=False)
    s = g.last_stats[0]
    rows.append((f, s))
    print(f"{f:4s} lines={s['n_active']:6d} evals={s['evals']:.3e} scale={s['scale_ms']:.3f} sum={s['sum_ms']:.3f} "
          f"fixup={s['fixup_ms']:.3f} ped={s['pedestal_ms']:.3f} total={s['total_ms']:.3f} ms  "
          f"sum rate={s['evals']/s['sum_ms']/1e9:.1f} Gevals/s")
tot = {k: sum(s[k] for _, s in rows) for k in ("scale_ms", "sum_ms", "fixup_ms", "pedestal_ms", "total_ms", "evals")}
print("TOTAL", {k: round(v, 3) if k != "evals" else v for k, v in tot.items()})
