#!/usr/bin/env python
"""Benchmark of the line-by-line hot path (BASELINE.json metric: Voigt line-grid evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one clear-sky column of BASELINE.json configs[1]: 60 layers x 7 gases
(H2O, CO2, O3, N2O, CO, CH4, O2; ~350 000 synthetic HITRAN-shaped lines) on the grid
1-5000 cm-1 at 0.01 cm-1 (v0=1, vn=5001, n_per_v=100; 500 000 points per spectrum), with
``remove_pedestal=True`` (what ``Spectroscopy.compute_absorption`` passes by default,
pyLBL/spectroscopy.py:163-164) and ``cut_off=25``.  With N ranks, rank r computes its own
column r (seeded perturbation of the standard column): weak scaling, no collective on the
data path.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value      line-grid evaluations / s with the packed line lists resident in HBM and the
             spectra left in HBM; device-timed (CUDA events), max over ranks
  e2e        the same metric through the public ``Gas`` API with HOST buffers: per step the
             (p, T, vmr) arrays go host->device and every spectrum comes back to pinned host
             memory inside the timed region
  roofline   FP64-pipe roofline of the summation kernel: algorithmic flops (7.3 per
             evaluation) / CUDA-event duration of its launches,
             against the FP64 FMA peak measured live on this GPU
  cpu_baseline  the reference C library (oracle/_ref, or the oracle port when absent) on
             the host cores, on a bounded sample of the same workload
"""
from __future__ import annotations

import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA context: see pylbl_b200/_lib.py
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from pylbl_b200 import synth  # noqa: E402

METRIC = "voigt_line_grid_evals_per_s"
UNIT = "evals/s"
CONFIG = 2
FLOP_PER_EVAL = 7.0  # algorithmic flops of one Lorentz-form evaluation
                     # (the only kind the summation kernel performs; regions 1-3/CPF12 are K2b's)
GASES = ["H2O", "CO2", "O3", "N2O", "CO", "CH4", "O2"]
# Order in which a step submits the gases: fewest lines first.  Every gas returns the same
# number of bytes, so the short ones go first and their copies to the host are under way
# while the long ones compute (two-stage flow shop, Johnson's rule).
SUBMIT_ORDER = sorted(GASES, key=lambda f: synth.CONFIG2_SHARES[f])
if os.environ.get("BENCH_GAS_ORDER"):
    SUBMIT_ORDER = os.environ["BENCH_GAS_ORDER"].split(",")
N_LAYERS = 60
REMOVE_PEDESTAL = os.environ.get("BENCH_PEDESTAL", "1") != "0"   # BASELINE: pedestal on (the knob is for experiments)
CUT_OFF = 25


def workload_config(n_gpus):
    v0, vn, npv = synth.config_grid(CONFIG)
    return {
        "workload": "BASELINE configs[1]: 60-layer clear-sky column x 7 gases "
                    "(H2O,CO2,O3,N2O,CO,CH4,O2), ~350k synthetic HITRAN-shaped lines, "
                    "grid 1-5000 cm-1 @0.01 cm-1",
        "v0": v0, "vn": vn, "n_per_v": npv, "n_points": (vn - v0) * npv,
        "n_layers": N_LAYERS, "gases": GASES,
        "n_lines": int(sum(synth.CONFIG2_SHARES.values())),
        "remove_pedestal": REMOVE_PEDESTAL, "cut_off": CUT_OFF,
        "columns_per_step": n_gpus, "sharding": f"one column per GPU x{n_gpus}, no collective",
        "l2_policy": "inputs larger than L2: per step ~1.5 GB of scaled line records are "
                     "rewritten and ~1.7 GB of spectra written (L2 = 126 MB)",
    }


def database_path(local_rank, barrier):
    """One synthetic database per node, written by local rank 0."""
    cache = Path(tempfile.gettempdir()) / "pylbl_b200_bench"
    cache.mkdir(exist_ok=True)
    path = cache / f"config{CONFIG}.db"
    done = cache / f"config{CONFIG}.done"
    if local_rank == 0 and not done.exists():
        synth.write_database(str(path), synth.config_line_lists(CONFIG))
        done.write_text("ok")
    barrier()
    while not done.exists():
        time.sleep(0.2)
    return str(path)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []      # (arrival time, csv line)
        self.window = None   # (begin, end) of the timed region, time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.window = (time.perf_counter(), None)

    def mark_end(self):
        self.window = (self.window[0], time.perf_counter())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # nvidia-smi is started well before the timed region (attaching to the driver stalls the
        # GPU for tens of milliseconds); only the samples that arrived during the region count.
        lo, hi = self.window if self.window and self.window[1] else (0., float("inf"))
        inside = [text for (t, text) in self.lines if lo <= t <= hi + 0.12]
        for line in (inside or [text for (_, text) in self.lines[-3:]]):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing more than half of the highest power seen
        loaded = [s for s, w in zip(sm, power) if power and w >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded) if loaded else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm: the reference C library on the host cores
# --------------------------------------------------------------------------------------
def cpu_reference_run(db, column, layers, threads):
    """Runs absorption() for every gas x the given layers on `threads` host threads (ctypes
    releases the GIL; the library is stateless).  Returns
    (evals, seconds, kind)."""
    from oracle import OracleGas, ReferenceGas, have_reference
    v0, vn, npv = synth.config_grid(CONFIG)
    kind = "reference" if have_reference() else "port"
    counters = {f: OracleGas(db, f) for f in GASES} if kind == "port" else None
    jobs = [(f, l) for l in layers for f in GASES]

    def one(job):
        f, l = job
        if kind == "reference":
            ReferenceGas(db, f).absorption(column.t[l], column.p[l], column.vmr[f][l], v0, vn, npv,
                                           REMOVE_PEDESTAL, CUT_OFF)
        else:
            counters[f].absorption(column.t[l], column.p[l], column.vmr[f][l], v0, vn, npv,
                                   REMOVE_PEDESTAL, CUT_OFF)
        return 0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(one, jobs))
    seconds = time.perf_counter() - t0
    return seconds, kind, len(jobs)


def count_evals(db, column, layers):
    """Evaluation count of gas x layers from the window arithmetic alone (no Voigt work):
    sum over lines of (e - s + 1), spectra.c:48-62, via the oracle's window function."""
    from oracle import read_molecule
    v0, vn, npv = synth.config_grid(CONFIG)
    n = (vn - v0) * npv
    total = 0
    for f in GASES:
        d = read_molecule(db, f)
        nu, delta = d["nu"], d["delta_air"]
        stop = np.nonzero((nu > vn + CUT_OFF + 1) | (nu < v0 - (CUT_OFF + 1)))[0]
        na = int(stop[0]) if stop.size else nu.size
        for l in layers:
            p_atm = column.p[l] * 9.86923e-6
            cb = np.floor(nu[:na] + p_atm * delta[:na]) - v0
            s = (cb - CUT_OFF) * npv
            e = np.minimum((cb + CUT_OFF + 1) * npv, n - 1)
            keep = s < n
            s = np.maximum(s, 0)
            total += int(np.sum((e - s + 1)[keep & (e >= s)]))
    return total


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    db = database_path(0, lambda: None)
    column = synth.standard_column(N_LAYERS, column=0)
    threads = os.cpu_count() or 1
    layers = [0, 20, 40, 59]
    evals = count_evals(db, column, layers)
    times = []
    kind = "port"
    for step in range(args.warmup + args.steps):
        seconds, kind, jobs = cpu_reference_run(db, column, layers, threads)
        if step >= args.warmup:
            times.append(seconds)
    mean = sum(times) / len(times)
    value = evals / mean
    sample = (f"{len(GASES)} gases x layers {layers} of the 60-layer column "
              f"({len(GASES) * len(layers)} absorption() calls, {evals:.3e} evaluations per step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "layer_spectra_per_s": len(layers) / mean,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# rank plumbing (the data path has no collective: ranks only agree on timing and totals)
# --------------------------------------------------------------------------------------
class Ranks(object):
    """barrier / max / sum over the ranks of a torch.distributed job (or a single process)."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, None if self.dist is None else self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, None if self.dist is None else self.dist.ReduceOp.SUM)


def rank_column(rank):
    """Weak scaling: rank r computes its own 60-layer column r."""
    return synth.standard_column(N_LAYERS, column=rank)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world, dist):
    import torch
    from pylbl_b200 import Gas, _lib

    torch.cuda.set_device(local_rank)
    lib = _lib.library()
    ranks = Ranks(dist, f"cuda:{local_rank}")
    barrier, max_over_ranks, sum_over_ranks = ranks.barrier, ranks.max, ranks.sum

    db = database_path(local_rank, barrier)
    bounds = synth.config_grid(CONFIG)
    v0, vn, npv = bounds
    n = (vn - v0) * npv
    column = rank_column(rank)
    gases = {f: Gas(db, f, devices=[local_rank]) for f in GASES}

    peak = ctypes.c_double(0.)
    lib.lbl_measure_fp64_peak(local_rank, ctypes.byref(peak))

    def submit_all(destinations):
        """Submits every gas of the column (each handle has its own CUDA streams, the
        summation kernels share the device's main stream in submission order), then waits."""
        handles = []
        for f in SUBMIT_ORDER:
            h = gases[f]._handle(local_rank)
            t = np.ascontiguousarray(column.t)
            p = np.ascontiguousarray(column.p)
            x = np.ascontiguousarray(column.vmr[f])
            dst = destinations[f].array.ctypes.data_as(ctypes.c_void_p) if destinations else None
            lib.lbl_gas_submit(h.ptr, N_LAYERS, p, t, x, v0, vn, npv, CUT_OFF,
                               1 if REMOVE_PEDESTAL else 0, 0, dst)
            handles.append(h)
        stats = []
        for h in handles:
            lib.lbl_gas_wait(h.ptr)
            stats.append(h.stats())
        return stats

    def step_resident():
        """All gases, spectra left on the device."""
        return submit_all(None)

    pinned = {f: _lib.PinnedArray((N_LAYERS, n)) for f in GASES}

    def step_e2e():
        """Public API with host buffers: inputs go up and every spectrum comes back to pinned
        host memory."""
        return submit_all(pinned)

    # ---- device-resident throughput ("value") ------------------------------------------
    # Settle first: on a fresh box the first steps run slow (allocations, clocks and power
    # state ramping up); repeat untimed steps until two in a row agree within 3 %, at most 20.
    # The W warm-up steps asked for come after that.
    sampler = ClockSampler(local_rank)
    sampler.start()
    previous = None
    for _ in range(20):
        lib.lbl_timer_start(local_rank)
        step_resident()
        t_step = ctypes.c_float(0.)
        lib.lbl_timer_stop(local_rank, ctypes.byref(t_step))
        settled = previous is not None and abs(t_step.value - previous) <= 0.03 * previous
        previous = t_step.value
        if os.environ.get("BENCH_DEBUG"):
            print(f"settle step: {t_step.value:.2f} ms", file=sys.stderr)
        if ranks.max(0.0 if settled else 1.0) == 0.0:
            break
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    barrier()
    sampler.mark_begin()
    lib.lbl_timer_start(local_rank)
    evals = 0
    executed = 0
    sum_ms = 0.0
    sum_launches = 0
    launches = 0
    points = cells = 0
    roof_evals = 0
    FARFIELD_GRID = npv >= 64
    for _ in range(args.steps):
        if os.environ.get("BENCH_DEBUG"):
            _t0 = time.perf_counter()
            _st = step_resident()
            print(f"timed step: {(time.perf_counter() - _t0) * 1e3:.2f} ms wall", file=sys.stderr)
        else:
            _st = step_resident()
        for s in _st:
            evals += s["evals"]
            launches += s["total_launches"]
            # The roofline is that of the dominant kernel: only the gases summed by the far-field
            # kernel K2c count for it (gases with very few lines go through the direct kernel).
            if s["cells_per_warp"] or not FARFIELD_GRID:
                points = s["points_per_thread"]
                cells = s["cells_per_warp"]
                executed += s["executed"]
                roof_evals += s["evals"]
                sum_ms += s["sum_ms"]
                sum_launches += s["sum_launches"]
    ms = ctypes.c_float(0.)
    lib.lbl_timer_stop(local_rank, ctypes.byref(ms))
    sampler.mark_end()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    seconds = max_over_ranks(ms.value * 1e-3)
    total_evals = sum_over_ranks(float(evals))
    value = total_evals / seconds

    # ---- end to end through the public API with host buffers ---------------------------
    for _ in range(args.warmup):
        step_e2e()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    e2e_evals = 0
    h2d = d2h = 0
    for _ in range(args.steps):
        _t0 = time.perf_counter()
        if os.environ.get("BENCH_DEBUG"):
            lib.lbl_timer_start(local_rank)   # origin of PYLBL_B200_TIMELINE's printout
        _st = step_e2e()
        if os.environ.get("BENCH_DEBUG"):
            print(f"e2e step: {(time.perf_counter() - _t0) * 1e3:.2f} ms wall", file=sys.stderr)
        for s in _st:
            e2e_evals += s["evals"]
            h2d += s["h2d_bytes"]
            d2h += s["d2h_bytes"]
    torch.cuda.synchronize()
    e2e_seconds = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = sum_over_ranks(float(e2e_evals)) / e2e_seconds
    h2d = int(sum_over_ranks(float(h2d)))     # whole job, like `value`
    d2h = int(sum_over_ranks(float(d2h)))
    launches = int(sum_over_ranks(float(launches)))
    # ---- the summation kernel alone (untimed extra pass) ---------------------------------
    # One more pass with the gases run one at a time and the pedestal off (the summation kernel
    # does not depend on it): nothing else is on the GPU while the kernel runs, which gives
    # its own duration.  In the step above the launches of the first gases share the SMs with
    # the scaling and pedestal kernels of all the gases of the column.
    isolated_ms = 0.0
    isolated_launches = 0
    for f in SUBMIT_ORDER:
        gases[f].absorption_coefficients(column.t, column.p, column.vmr[f], bounds=bounds,
                                         remove_pedestal=False, cut_off=CUT_OFF, to_host=False)
        st = gases[f].last_stats[0]
        if st["cells_per_warp"] or not FARFIELD_GRID:
            isolated_ms += st["sum_ms"]
            isolated_launches += st["sum_launches"]

    # parity spot check of what came back (one spectrum, against the oracle)
    check = None
    if rank == 0 and not args.no_check:
        from oracle import OracleGas
        ref = OracleGas(db, "CO")
        layer = 37
        k_ref = ref.absorption(column.t[layer], column.p[layer], column.vmr["CO"][layer], v0, vn,
                               npv, REMOVE_PEDESTAL, CUT_OFF)
        check = float(np.max(np.abs(pinned["CO"].array[layer] - k_ref)) / np.max(np.abs(k_ref)))

    # ---- roofline of the summation kernel ----------------------------------------------
    # The summation kernel on fine grids interpolates the far field: it PERFORMS `executed`
    # Lorentz evaluations to deliver `evals` reference-equivalent ones.  The roofline counts
    # the work performed; `value` counts the work delivered.
    # K2c also evaluates the interpolant at every point: Clenshaw over the summed series, one
    # FMA + one add per (point, coefficient), 32 coefficients; and 2*(32^2 + 16^2 + 8^2) flop
    # per cell for the node-sum -> coefficient transforms.
    interp_flops = ((3.0 * 32 * n + 2.0 * (32 * 32 + 16 * 16 + 8 * 8) * (vn - v0))
                    * N_LAYERS * sum_launches) if cells else 0.0
    flops = FLOP_PER_EVAL * executed + interp_flops     # this rank, timed region
    achieved = flops / (sum_ms * 1e-3) / 1e12 if sum_ms > 0 else 0.0
    kernel = f"lbl::sum_cell_kernel<{cells}>" if cells else f"lbl::sum_kernel<{points}>"
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
    # (profiles/r1_sum_cell_dram.json, written by tools/ncu_dram.py), or null.
    traffic = None
    dram_file = ROOT / "profiles" / "r1_sum_cell_dram.json"
    if cells and dram_file.exists():
        traffic = json.loads(dram_file.read_text()).get("dram_bytes_per_launch")
    roofline = {
        "bound": "fp64", "kernel": kernel, "achieved": achieved,
        "peak": peak.value, "unit": "TFLOP/s",
        "frac": achieved / peak.value if peak.value else None, "traffic": traffic,
        "flop_per_eval": FLOP_PER_EVAL,
        # SURVEY 8(d)'s figure: 7.3 flop x reference-equivalent evaluations DELIVERED by the
        # launches / their time.  Above the FP64 peak because the polynomial far field
        # delivers several evaluations per evaluation performed (far_field_work_reduction);
        # `achieved`/`frac` above count only the arithmetic the kernel really executes.
        "achieved_reference_equivalent": 7.3 * roof_evals / (sum_ms * 1e-3) / 1e12 if sum_ms > 0 else None,
        "frac_reference_equivalent": (7.3 * roof_evals / (sum_ms * 1e-3) / 1e12) / peak.value
        if sum_ms > 0 and peak.value else None,
        "interpolation_flops_per_launch": interp_flops / max(sum_launches, 1),
        "executed_evals_per_launch": executed / max(sum_launches, 1),
        "reference_evals_per_launch": roof_evals / max(sum_launches, 1),
        "far_field_work_reduction": roof_evals / executed if executed else None,
        "launches_counted": sum_launches,
        "avg_launch_ms": sum_ms / max(sum_launches, 1),
        # the same launches with nothing else on the GPU (one gas at a time, pedestal off)
        "isolated_avg_launch_ms": isolated_ms / max(isolated_launches, 1),
        "frac_isolated": (flops / args.steps / (isolated_ms * 1e-3) / 1e12) / peak.value
        if isolated_ms > 0 and peak.value else None,
        "kernel_share_of_step": sum_ms / (ms.value if ms.value else 1.0),
        "peak_source": "FP64 FMA peak measured live on this GPU (independent DFMA chains, "
                       "lbl_measure_fp64_peak); MEASURED_PEAKS.json carries no FP64 figure",
        "hbm_peak_gbs": json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs")
        if (ROOT / "MEASURED_PEAKS.json").exists() else None,
    }

    # ---- CPU baseline on rank 0 at N=1 ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        layers = [0, 20, 40, 59]
        base_column = synth.standard_column(N_LAYERS, column=0)
        cpu_evals = count_evals(db, base_column, layers)
        cpu_seconds, kind, jobs = cpu_reference_run(db, base_column, layers, threads)
        cpu = {"value": cpu_evals / cpu_seconds, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{len(GASES)} gases x layers {layers} of the 60-layer column "
                         f"({jobs} absorption() calls incl. their sqlite reads, "
                         f"{cpu_evals:.3e} evaluations, {cpu_seconds:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": seconds * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "layer_spectra_per_s": world * N_LAYERS * args.steps / seconds,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": e2e_seconds * 1e3 / args.steps,
                    "layer_spectra_per_s": world * N_LAYERS * args.steps / e2e_seconds,
                    "parity_spot_check_max_rel_to_peak": check},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    for g in gases.values():
        g.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist = dist_mod
    try:
        run_ours(args, rank, local_rank, world, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
