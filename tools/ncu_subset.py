"""The metrics quoted in DESIGN.md / profiles/README.md from `ncu --page raw --csv` dumps:
python tools/ncu_subset.py OUT.csv RAW.csv [RAW.csv ...]"""
import csv
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "not_selected",
          "no_instruction", "branch_resolving", "mio_throttle", "lg_throttle", "membar", "dispatch_stall",
          "drain", "imc_miss", "tex_throttle", "sleeping", "selected"]
WANT += [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALLS]

out, paths = sys.argv[1], sys.argv[2:]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    first = True
    for path in paths:
        rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
        head, units = rows[0], rows[1]
        idx = [head.index(k) for k in WANT if k in head]
        if first:
            w.writerow(["source"] + [head[i] for i in idx])
            w.writerow([""] + [units[i] for i in idx])
            first = False
        for r in rows[2:]:
            if len(r) == len(head):
                w.writerow([path.split("/")[-1]] + [r[i] for i in idx])
print(out)
