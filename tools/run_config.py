"""Runs one BASELINE.json configuration through the Gas API and prints timing per gas.
   python tools/run_config.py CONFIG [n_layers] [--no-pedestal] [--fp32]"""
import sys, time, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pylbl_b200 import Gas, synth

config = int(sys.argv[1])
n_layers = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 60
ped = "--no-pedestal" not in sys.argv
prec = "fp32" if "--fp32" in sys.argv else "fp64"
cache = Path(tempfile.gettempdir()) / "pylbl_b200_bench"
cache.mkdir(exist_ok=True)
db = cache / f"config{config}.db"
t0 = time.time()
lists = synth.config_line_lists(2 if config == 5 else config)
if not db.exists():
    synth.write_database(str(db), lists)
print(f"database ready in {time.time()-t0:.1f} s")
bounds = synth.config_grid(config)
col = synth.standard_column(n_layers)
tot_ms = 0.0; tot_ev = 0
for f in lists:
    t0 = time.time(); g = Gas(str(db), f, devices=[0], precision=prec)
    g.absorption_coefficients(col.t[:1], col.p[:1], col.vmr[f][:1], bounds=bounds, remove_pedestal=ped, to_host=False)
    t_open = time.time() - t0
    for rep in range(2):
        g.absorption_coefficients(col.t, col.p, col.vmr[f], bounds=bounds, remove_pedestal=ped, to_host=False)
    s = g.last_stats[0]
    tot_ms += s["total_ms"]; tot_ev += s["evals"]
    print(f"{f:4s} open+first={t_open:.2f}s lines={s['n_active']} P={s['points_per_thread']} evals={s['evals']:.3e} "
          f"scale={s['scale_ms']:.2f} sum={s['sum_ms']:.2f} fixup={s['fixup_ms']:.2f} ped={s['pedestal_ms']:.2f} "
          f"total={s['total_ms']:.2f} ms launches={s['total_launches']} -> {s['evals']/s['total_ms']/1e9:.3f} Tevals/s")
    g.close()
print(f"config {config}: {tot_ev:.3e} evals in {tot_ms:.1f} ms = {tot_ev/tot_ms/1e9:.3f} Tevals/s")
