"""One gas of the bench column (for ncu captures): python tools/one_gas.py [FORMULA] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pylbl_b200 import Gas, synth

f = sys.argv[1] if len(sys.argv) > 1 else "H2O"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ped = "--no-pedestal" not in sys.argv
db = bench.database_path(0, lambda: None)
bounds = synth.config_grid(5 if "--config5" in sys.argv else 2)
col = synth.standard_column(60)
g = Gas(db, f, devices=[0])
for rep in range(reps):
    g.absorption_coefficients(col.t, col.p, col.vmr[f], bounds=bounds, remove_pedestal=ped, to_host=False)
s = g.last_stats[0]
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()})
