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
  e2e        the same metric through the public API with HOST buffers, gas-summed output
             (``Mixture.total_absorption``, pyLBL's output_format="total"): per step the
             (p, T, vmr) arrays go host->device, the gases are scaled by their number densities
             and summed on the device, and ONE array per column comes back to pinned host memory
             inside the timed region
  e2e_gas    the same with per-gas output (``Gas.submit``/``Gas.wait``; output_format="gas"):
             seven arrays per column come back.  Both carry the measured device-to-host copy
             ceiling for their bytes (all ranks copying at once) and the step's floor
             max(kernels, copy)
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
# Layer groups of the gas-summed end-to-end path (Mixture.total_absorption): a finished group's
# rows travel to the host while the next group computes.
LAYER_GROUPS = int(os.environ["BENCH_LAYER_GROUPS"]) if os.environ.get("BENCH_LAYER_GROUPS") else None


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
def cpu_sample_layers(threads):
    """Layers of the 60-layer column that make up the CPU sample: enough (gas, layer) jobs to
    keep every host thread busy (>= 4 jobs per thread, spread evenly over the column)."""
    count = min(N_LAYERS, max(4, -(-4 * threads // len(GASES))))
    return sorted(set(int(round(i)) for i in np.linspace(0, N_LAYERS - 1, count)))


def cpu_reference_run(db, column, layers, threads):
    """Runs absorption() for every gas x the given layers on `threads` host threads (ctypes
    releases the GIL; the library is stateless); the longest jobs are
    queued first.  threads == 1 is the reference's own serial gas-outer / layer-inner loop
    (pyLBL/spectroscopy.py:166-191).  Returns (seconds, kind, jobs)."""
    from oracle import OracleGas, ReferenceGas, have_reference
    v0, vn, npv = synth.config_grid(CONFIG)
    kind = "reference" if have_reference() else "port"
    counters = {f: OracleGas(db, f) for f in GASES} if kind == "port" else None
    jobs = [(f, l) for f in sorted(GASES, key=lambda f: -synth.CONFIG2_SHARES[f]) for l in layers]

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
    if threads == 1:
        for job in jobs:
            one(job)
    else:
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
    layers = cpu_sample_layers(threads)
    evals = count_evals(db, column, layers)
    times = []
    kind = "port"
    for step in range(args.warmup + args.steps):
        seconds, kind, jobs = cpu_reference_run(db, column, layers, threads)
        if step >= args.warmup:
            times.append(seconds)
    mean = sum(times) / len(times)
    value = evals / mean
    sample = (f"{len(GASES)} gases x {len(layers)} layers {layers} of the 60-layer column "
              f"({len(GASES) * len(layers)} absorption() calls incl. their sqlite reads, longest first, "
              f"on {threads} threads; {evals:.3e} evaluations per step)")
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


def rank_columns(n_columns, world, rank):
    """configs[4]: the columns of the batch dealt round-robin to the ranks."""
    return [c for c in range(n_columns) if c % world == rank]


def rank_band(edges, rank):
    """configs[3]: rank r takes cells [edges[r], edges[r+1]) of the one grid."""
    return int(edges[rank]), int(edges[rank + 1])


def rank_column(rank):
    """Weak scaling: rank r computes its own 60-layer column r."""
    return synth.standard_column(N_LAYERS, column=rank)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Keeps this rank's host threads -- and with them the page-locked buffers it allocates
    (first touch) -- on the NUMA node its GPU hangs off: a device-to-host copy that crosses the
    socket interconnect is the first thing that slows down when eight GPUs copy at once.
    Returns what was done (for the JSON line); any failure leaves the process unbound."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        name = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{name}/numa_node").read_text())
        if node < 0:
            return {"node": None, "why": "no NUMA information for the GPU"}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "why": "none of the node's CPUs is available to this process"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as exc:   # no sysfs, no permission, ...: run unbound
        return {"node": None, "why": f"{type(exc).__name__}: {exc}"}


def copy_ceiling(local_rank, nbytes, barrier, max_over_ranks, repeats=5):
    """Plain device-to-host copy of `nbytes` into page-locked memory on every rank at once:
    the ceiling of any end-to-end number that returns that many bytes.  GB/s of this rank's
    copy, slowest rank."""
    import torch
    dev = torch.empty(nbytes // 8, dtype=torch.float64, device=f"cuda:{local_rank}")
    host = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(repeats):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host.copy_(dev, non_blocking=True)
        e1.record()
        e1.synchronize()
        seconds = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        best = max(best, nbytes / seconds / 1e9)
    del dev, host
    return best


def run_ours(args, rank, local_rank, world, dist):
    import torch
    from pylbl_b200 import Gas, Mixture, _lib

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if os.environ.get("BENCH_NUMA", "1") != "0" else None
    lib = _lib.library()
    ranks = Ranks(dist, f"cuda:{local_rank}")
    barrier, max_over_ranks, sum_over_ranks = ranks.barrier, ranks.max, ranks.sum

    db = database_path(local_rank, barrier)
    bounds = synth.config_grid(CONFIG)
    v0, vn, npv = bounds
    n = (vn - v0) * npv
    column = rank_column(rank)
    gases = {f: Gas(db, f, devices=[local_rank]) for f in GASES}
    mixture = Mixture.from_gases({f: gases[f] for f in SUBMIT_ORDER}, local_rank)

    peak = ctypes.c_double(0.)
    lib.lbl_measure_fp64_peak(local_rank, ctypes.byref(peak))

    def submit_all(destinations):
        """Submits every gas of the column through the public non-blocking `Gas.submit` (the
        gases overlap on the device: scaling kernels and pedestal chains side by side, summation
        kernels gas after gas, copies under the next gas's kernels), then waits for all."""
        for f in SUBMIT_ORDER:
            gases[f].submit(column.t, column.p, column.vmr[f], bounds=bounds,
                            remove_pedestal=REMOVE_PEDESTAL, cut_off=CUT_OFF,
                            out=destinations[f].array if destinations else None)
        return [gases[f].wait() for f in SUBMIT_ORDER]

    def step_resident():
        """All gases, spectra left on the device."""
        return submit_all(None)

    pinned = {f: _lib.PinnedArray((N_LAYERS, n)) for f in GASES}
    pinned_total = _lib.PinnedArray((N_LAYERS, n))

    def step_e2e_gas():
        """Per-gas output (pyLBL's output_format="gas"/"all": one array per gas): inputs go up
        and all seven spectra arrays come back to pinned host memory."""
        return submit_all(pinned)

    def step_e2e_total():
        """Gas-summed output (pyLBL's output_format="total", spectroscopy.py:225-234): the
        number densities are applied and the gases summed on the device; one array comes back."""
        mixture.total_absorption(column.t, column.p, column.vmr, bounds=bounds,
                                 remove_pedestal=REMOVE_PEDESTAL, cut_off=CUT_OFF,
                                 out=pinned_total.array, layer_groups=LAYER_GROUPS)
        return [gases[f].last_stats[0] for f in SUBMIT_ORDER]

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
    def time_e2e(step):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        n_evals = 0
        up = down = 0
        for _ in range(args.steps):
            _t0 = time.perf_counter()
            if os.environ.get("BENCH_DEBUG"):
                lib.lbl_timer_start(local_rank)   # origin of PYLBL_B200_TIMELINE's printout
            _st = step()
            if os.environ.get("BENCH_DEBUG"):
                print(f"e2e step: {(time.perf_counter() - _t0) * 1e3:.2f} ms wall", file=sys.stderr)
            for s in _st:
                n_evals += s["evals"]
                up += s["h2d_bytes"]
                down += s["d2h_bytes"]
        torch.cuda.synchronize()
        e2e_seconds = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return {"value": sum_over_ranks(float(n_evals)) / e2e_seconds, "unit": UNIT,
                "h2d_bytes_per_step": int(sum_over_ranks(float(up))) // args.steps,      # whole job,
                "d2h_bytes_per_step": int(sum_over_ranks(float(down))) // args.steps,    # like `value`
                "ms_per_step": e2e_seconds * 1e3 / args.steps,
                "layer_spectra_per_s": world * N_LAYERS * args.steps / e2e_seconds}

    e2e_total = time_e2e(step_e2e_total)
    e2e_total["layer_groups"] = mixture.last_layer_groups
    e2e_total["last_call_uncovered_copy_ms"] = mixture.last_copy_tail_ms
    e2e_total["output"] = ("total: sum over gases of n_gas*k_gas formed on the device, one "
                           "(60, 500000) f64 array per column back to pinned host memory "
                           "(pyLBL output_format='total'); public API Mixture.total_absorption")
    e2e_gas = time_e2e(step_e2e_gas)
    e2e_gas["output"] = ("gas: seven (60, 500000) f64 arrays per column back to pinned host memory "
                         "(pyLBL output_format='gas'/'all'); public API Gas.submit / Gas.wait")
    launches = int(sum_over_ranks(float(launches)))

    # ---- the copy ceiling: the same bytes per rank, copy alone, all ranks at once ----------
    ceiling_gas = copy_ceiling(local_rank, 8 * N_LAYERS * n * len(GASES), barrier, max_over_ranks)
    ceiling_total = copy_ceiling(local_rank, 8 * N_LAYERS * n, barrier, max_over_ranks)
    for block, gbs, nbytes in ((e2e_gas, ceiling_gas, 8 * N_LAYERS * n * len(GASES)),
                               (e2e_total, ceiling_total, 8 * N_LAYERS * n)):
        copy_ms = nbytes / (gbs * 1e9) * 1e3
        block["d2h_ceiling_gbs_per_gpu"] = gbs
        block["d2h_alone_ms_per_step"] = copy_ms
        # a step cannot be shorter than its kernels nor than its copy
        block["floor_ms_per_step"] = max(copy_ms, seconds * 1e3 / args.steps)
        block["frac_of_floor"] = block["floor_ms_per_step"] / block["ms_per_step"]

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

    # parity spot check of what came back, against the oracle: a gas the far-field kernel
    # summed (CO2, the longest line list), one layer; with the pedestal as benchmarked in the
    # window-scaled metric, and without it pointwise
    check = None
    if rank == 0 and not args.no_check:
        from oracle import OracleGas
        sys.path.insert(0, str(ROOT / "tests"))
        from helpers import relative_error, scaled_error
        ref = OracleGas(db, "CO2")
        layer = 37
        state = (column.t[layer], column.p[layer], column.vmr["CO2"][layer])
        k_ref = ref.absorption(*state, v0, vn, npv, REMOVE_PEDESTAL, CUT_OFF)
        check = {"gas": "CO2", "layer": layer,
                 "scaled_error_as_benchmarked": scaled_error(pinned["CO2"].array[layer], k_ref, npv, CUT_OFF)}
        k_plain = gases["CO2"].absorption_coefficients([state[0]], [state[1]], [state[2]], bounds=bounds,
                                                       remove_pedestal=False, cut_off=CUT_OFF)[0]
        check["cells_per_warp"] = gases["CO2"].last_stats[0]["cells_per_warp"]
        check["pointwise_relative_error_no_pedestal"] = relative_error(
            k_plain, ref.absorption(*state, v0, vn, npv, False, CUT_OFF))
        # the device-side sum against the per-gas arrays that came back
        from pylbl_b200 import number_density
        want = np.zeros(n)
        for f in GASES:
            want += number_density(*state[:2], column.vmr[f][layer]) * pinned[f].array[layer]
        check["total_vs_sum_of_gases_scaled_error"] = scaled_error(pinned_total.array[layer], want, npv, CUT_OFF)

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
    # (written by tools/ncu_dram.py), or null.
    traffic = None
    for name in ("r2_sum_cell_dram.json", "r1_sum_cell_dram.json"):
        dram_file = ROOT / "profiles" / name
        if cells and dram_file.exists():
            traffic = json.loads(dram_file.read_text()).get("dram_bytes_per_launch")
            break
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
    cpu = cpu_serial = cpu_config1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, range(os.cpu_count() or 1)) if numa and numa.get("cpus") else None
        threads = os.cpu_count() or 1
        layers = cpu_sample_layers(threads)
        base_column = synth.standard_column(N_LAYERS, column=0)
        cpu_evals = count_evals(db, base_column, layers)
        cpu_seconds, kind, jobs = cpu_reference_run(db, base_column, layers, threads)
        cpu = {"value": cpu_evals / cpu_seconds, "unit": UNIT, "cores": threads, "kind": kind,
               "mode": "(ii) all host cores: thread pool over (gas, layer) calls",
               "layer_spectra_per_s": len(layers) / cpu_seconds,
               "sample": f"{len(GASES)} gases x {len(layers)} layers of the 60-layer column "
                         f"({jobs} absorption() calls incl. their sqlite reads, longest first, "
                         f"{cpu_evals:.3e} evaluations, {cpu_seconds:.1f} s)"}
        # mode (i), "as shipped": one thread, the reference driver's serial gas-outer /
        # layer-inner loop (pyLBL/spectroscopy.py:166-191), sqlite re-read on every call
        serial_layers = [0, N_LAYERS - 1]
        serial_evals = count_evals(db, base_column, serial_layers)
        serial_seconds, kind, jobs = cpu_reference_run(db, base_column, serial_layers, 1)
        cpu_serial = {"value": serial_evals / serial_seconds, "unit": UNIT, "cores": 1, "kind": kind,
                      "mode": "(i) as shipped: one thread, serial gas x layer loop",
                      "layer_spectra_per_s": len(serial_layers) / serial_seconds,
                      "sample": f"{len(GASES)} gases x layers {serial_layers} ({jobs} calls, "
                                f"{serial_evals:.3e} evaluations, {serial_seconds:.1f} s)"}
        cpu_config1 = cpu_config1_point(local_rank)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": seconds * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "layer_spectra_per_s": world * N_LAYERS * args.steps / seconds,
            "clocks": clocks,
            "e2e": e2e_total,
            "e2e_gas": e2e_gas,
            "parity_spot_check": check,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cpu_baseline_serial": cpu_serial,
            "cpu_config1": cpu_config1,
            "numa": numa,
        }
        print(json.dumps(line), flush=True)
    mixture.close()
    for g in gases.values():
        g.close()


def cpu_config1_point(device):
    """BASELINE configs[0], the reference's own CPU-runnable case: one layer (98388 Pa, 289 K),
    H2O+CO2+O3, ~50k lines, 1-5000 cm-1 @0.1: the reference on one host thread and this
    library (direct kernel K2) on the same database, pedestal as benchmarked."""
    from oracle import ReferenceGas, OracleGas, have_reference
    from pylbl_b200 import Gas
    cache = Path(tempfile.gettempdir()) / "pylbl_b200_bench"
    path = cache / "config1.db"
    if not path.exists():
        synth.write_database(str(path), synth.config_line_lists(1))
    atm = synth.fixture_atmosphere()
    layer = 3
    bounds = synth.config_grid(1)
    cls = ReferenceGas if have_reference() else OracleGas
    counter = {f: OracleGas(str(path), f) for f in ("H2O", "CO2", "O3")}
    evals = 0
    for f, ref in counter.items():
        ref.absorption(atm.t[layer], atm.p[layer], atm.vmr[f][layer], *bounds, REMOVE_PEDESTAL, CUT_OFF)
        evals += ref.last_evals
    best = float("inf")
    for _ in range(5):
        t0 = time.perf_counter()
        for f in counter:
            cls(str(path), f).absorption(atm.t[layer], atm.p[layer], atm.vmr[f][layer], *bounds,
                                         REMOVE_PEDESTAL, CUT_OFF)
        best = min(best, time.perf_counter() - t0)
    gases = {f: Gas(str(path), f, devices=[device]) for f in counter}
    gpu_best = float("inf")
    for _ in range(10):
        t0 = time.perf_counter()
        for f, g in gases.items():
            g.absorption_coefficient(atm.t[layer], atm.p[layer], atm.vmr[f][layer],
                                     synth.grid_from_bounds(*bounds), remove_pedestal=REMOVE_PEDESTAL,
                                     cut_off=CUT_OFF)
        gpu_best = min(gpu_best, time.perf_counter() - t0)
    for g in gases.values():
        g.close()
    return {"workload": "BASELINE configs[0]: one layer, H2O+CO2+O3 (~50k lines), 1-5000 cm-1 @0.1",
            "evals": evals, "cpu_value": evals / best, "cpu_seconds": best, "cpu_cores": 1,
            "cpu_kind": "reference" if have_reference() else "port",
            "gpu_value_scalar_plugin_calls": evals / gpu_best, "gpu_seconds": gpu_best, "unit": UNIT}


# --------------------------------------------------------------------------------------
# the other BASELINE configurations (run by hand; the driver runs the default, configs[1])
# --------------------------------------------------------------------------------------
def aux_database(config, local_rank, barrier):
    cache = Path(tempfile.gettempdir()) / "pylbl_b200_bench"
    cache.mkdir(exist_ok=True)
    path = cache / f"config{config}.db"
    done = cache / f"config{config}.done"
    if local_rank == 0 and not done.exists():
        synth.write_database(str(path), synth.config_line_lists(config))
        done.write_text("ok")
    barrier()
    while not done.exists():
        time.sleep(0.2)
    return str(path)


def timed_steps(args, step, barrier, max_over_ranks):
    """W warm-up steps, K timed ones; wall clock bracketed by synchronise + barrier, max over ranks."""
    import torch
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    seconds = max_over_ranks(time.perf_counter() - t0)
    barrier()
    return seconds


def run_config4(args, rank, local_rank, world, dist):
    """BASELINE configs[3]: ~1M-line list x 60 layers, 10-3500 cm-1 @0.001, one spectral band per
    GPU (strong scaling: the bands of the N ranks tile the one grid; no collective)."""
    import torch
    from pylbl_b200 import Gas, _lib
    torch.cuda.set_device(local_rank)
    ranks = Ranks(dist, f"cuda:{local_rank}")
    db = aux_database(4, local_rank, ranks.barrier)
    bounds = synth.config_grid(4)
    v0, vn, npv = bounds
    column = synth.standard_column(N_LAYERS, column=0)
    gas = Gas(db, "XX", devices=[local_rank])
    edges = gas.band_edges(bounds, world, CUT_OFF)
    lo, hi = rank_band(edges, rank)
    width = (hi - lo) * npv
    pinned = _lib.PinnedArray((N_LAYERS, width))
    lib = _lib.library()
    h = gas._handle(local_rank)

    def submit(dst):
        lib.lbl_gas_submit_band(h.ptr, N_LAYERS, np.ascontiguousarray(column.p), np.ascontiguousarray(column.t),
                                np.ascontiguousarray(column.vmr["XX"]), v0, vn, npv, CUT_OFF,
                                1 if REMOVE_PEDESTAL else 0, 0, lo, hi, dst, 0)
        lib.lbl_gas_wait(h.ptr)
        return h.stats()

    resident_seconds = timed_steps(args, lambda: submit(None), ranks.barrier, ranks.max)
    stats = h.stats()
    e2e_seconds = timed_steps(args, lambda: submit(pinned.array.ctypes.data_as(ctypes.c_void_p)),
                              ranks.barrier, ranks.max)
    # every reduction is a collective: all of them happen here, on every rank
    evals = ranks.sum(float(stats["evals"]))
    points = ranks.sum(float(width))
    slowest_kernels = ranks.max(stats["sum_ms"] + stats["fixup_ms"])
    h2d_total = int(ranks.sum(float(stats["h2d_bytes"])))
    launches_total = int(ranks.sum(float(stats["total_launches"])))
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": evals * args.steps / resident_seconds, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": resident_seconds * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: synthetic ~1M-line list x 60 layers, grid 10-3500 cm-1 "
                                   "@0.001 cm-1, one spectral band per GPU",
                       "v0": v0, "vn": vn, "n_per_v": npv, "n_points": int(points), "n_layers": N_LAYERS,
                       "n_lines": 1000000, "remove_pedestal": REMOVE_PEDESTAL, "cut_off": CUT_OFF,
                       "band_edges_cells": [int(x) for x in edges],
                       "sharding": f"{world} contiguous bands of about equal work (lbl_gas_band_edges), "
                                   "windows/prefix/pedestal of the whole grid, no collective"},
            "layer_spectra_per_s": N_LAYERS * args.steps / resident_seconds,
            "e2e": {"value": evals * args.steps / e2e_seconds, "unit": UNIT,
                    "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": int(8 * N_LAYERS * points), "ms_per_step": e2e_seconds * 1e3 / args.steps},
            "slowest_rank_sum_plus_near_ms": slowest_kernels, "gpu_launches": launches_total * args.steps,
        }), flush=True)
    gas.close()


def run_config5(args, rank, local_rank, world, dist):
    """BASELINE configs[4]: 256 columns x 60 layers, 7 gases + MT-CKD continuum on the grid
    1-5000 cm-1 @0.1, columns dealt round-robin to the GPUs, everything summed on
    the device: one (layers, 50000) array of total absorption per rank comes back."""
    import torch
    from pylbl_b200 import Continuum, Gas, Mixture, _lib
    torch.cuda.set_device(local_rank)
    ranks = Ranks(dist, f"cuda:{local_rank}")
    db = database_path(local_rank, ranks.barrier)          # the line lists of configs[1]
    bounds = synth.config_grid(5)
    v0, vn, npv = bounds
    n = (vn - v0) * npv
    n_columns = int(os.environ.get("BENCH_COLUMNS", "256"))
    mine = rank_columns(n_columns, world, rank)
    cols = [synth.standard_column(N_LAYERS, column=c) for c in mine]
    t = np.concatenate([c.t for c in cols])
    p = np.concatenate([c.p for c in cols])
    vmr = {f: np.concatenate([c.vmr[f] for c in cols]) for f in GASES}
    vmr["N2"] = np.full(t.size, 0.78)      # no lines in the database; the O2 and N2 continua read it
    gases = {f: Gas(db, f, devices=[local_rank]) for f in GASES}
    mixture = Mixture.from_gases(gases, local_rank)
    continuum = Continuum(local_rank)
    pinned = _lib.PinnedArray((t.size, n))

    def step():
        mixture.total_absorption(t, p, vmr, bounds=bounds, remove_pedestal=REMOVE_PEDESTAL, cut_off=CUT_OFF,
                                 out=pinned.array, continuum=continuum)

    seconds = timed_steps(args, step, ranks.barrier, ranks.max)
    stats = [gases[f].last_stats[0] for f in GASES]
    # every reduction is a collective: all of them happen here, on every rank
    evals = ranks.sum(float(sum(s["evals"] for s in stats)))
    kernels = ranks.max(sum(s["sum_ms"] + s["fixup_ms"] for s in stats))
    h2d_total = int(ranks.sum(float(sum(s["h2d_bytes"] for s in stats))))
    launches_total = int(ranks.sum(float(sum(s["total_launches"] for s in stats))))
    check = None
    if rank == 0 and not args.no_check:
        from oracle import OracleGas, mt_ckd
        from pylbl_b200 import continua_of, number_density
        sys.path.insert(0, str(ROOT / "tests"))
        from helpers import scaled_error
        row = 61 if t.size > 61 else 0                     # a layer of this rank's second column
        state = {f: float(vmr[f][row]) for f in vmr}
        grid = v0 + np.arange(n) * (1. / npv)
        want = np.zeros(n)
        for f in GASES:
            k = OracleGas(db, f).absorption(t[row], p[row], state[f], v0, vn, npv, REMOVE_PEDESTAL, CUT_OFF)
            want += number_density(t[row], p[row], state[f]) * k
        for f in vmr:
            for name in continua_of(f):
                want += mt_ckd.OracleContinuum(name).spectra(t[row], p[row], state, grid)
        check = {"row": row, "scaled_error_vs_oracle_lines_plus_continuum": scaled_error(pinned.array[row], want, npv, CUT_OFF)}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": evals * args.steps / seconds, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: {n_columns} columns x 60 layers, 7 gases + MT-CKD continuum, "
                                   "grid 1-5000 cm-1 @0.1 cm-1, columns round-robin over the GPUs",
                       "v0": v0, "vn": vn, "n_per_v": npv, "n_points": n, "n_layers": N_LAYERS,
                       "n_columns": n_columns, "gases": GASES + ["N2 (continuum only)"],
                       "remove_pedestal": REMOVE_PEDESTAL, "cut_off": CUT_OFF,
                       "output": "total absorption summed on the device, one array per rank to pinned host memory",
                       "timing": "end to end (host inputs, host output) -- the only number of this mode"},
            "layer_spectra_per_s": n_columns * N_LAYERS * args.steps / seconds,
            "e2e": {"value": evals * args.steps / seconds, "unit": UNIT,
                    "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": int(8 * n_columns * N_LAYERS * n), "ms_per_step": seconds * 1e3 / args.steps},
            "slowest_rank_sum_plus_near_ms_per_step": kernels,
            "direct_kernel": "lbl::sum_kernel<5>",
            "parity_spot_check": check,
            "gpu_launches": launches_total * args.steps,
        }), flush=True)
    mixture.close()
    continuum.close()
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
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="BASELINE configuration: 2 (the benchmark of record, default), "
                         "4 (1M lines, band-sharded: configs[3]), 5 (256 columns + continuum: configs[4])")
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
        import datetime
        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"),
                                    timeout=datetime.timedelta(seconds=180))
        dist = dist_mod
    try:
        if args.config == 4:
            run_config4(args, rank, local_rank, world, dist)
        elif args.config == 5:
            run_config5(args, rank, local_rank, world, dist)
        else:
            run_ours(args, rank, local_rank, world, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
