"""Handle open time, sqlite vs packed cache: python tools/open_time.py"""
import sys
import tempfile
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pylbl_b200 import Gas, synth
from pylbl_b200.gas_optics import cached_pack

db = bench.database_path(0, lambda: None)
cache = tempfile.mkdtemp(prefix="lblpack_")
col = synth.standard_column(4)
bounds = synth.config_grid(2)
Gas(db, "CO", devices=[0]).absorption_coefficients(col.t, col.p, col.vmr["CO"], bounds=bounds)  # CUDA init
for f in ("CO2", "H2O"):
    t0 = time.perf_counter()
    g = Gas(db, f, devices=[0]); g._handle(0)
    t_sql = time.perf_counter() - t0
    g.close()
    t0 = time.perf_counter(); cached_pack(db, f, cache); t_pack = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = Gas(db, f, devices=[0], cache_dir=cache); g._handle(0)
    t_open = time.perf_counter() - t0
    g.close()
    print(f"{f}: open from sqlite {t_sql*1e3:.1f} ms | write pack {t_pack*1e3:.1f} ms | open from pack {t_open*1e3:.1f} ms")
