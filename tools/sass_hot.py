"""Hot blocks of a kernel from an ncu report: python tools/sass_hot.py REPORT KERNEL_REGEX [min_pct]

Groups consecutive SASS instructions with the same execution count and prints each group's share
of executed warp instructions and of stall samples (needs a --set full --import-source capture)."""
import csv
import subprocess
import sys

rep, pattern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "-k", "regex:" + pattern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [n for n, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for k, s0 in enumerate(start):
    s1 = start[k + 1] if k + 1 < len(start) else len(rows)
    print("==", rows[s0][1][:100])
    h = rows[s0 + 1]
    data = [r for r in rows[s0 + 2:s1] if len(r) == len(h)]
    isrc, ismp, iex, ith = (h.index(x) for x in ("Source", "# Samples", "Instructions Executed",
                                                    "Avg. Threads Executed"))
    num = lambda x: int(x) if x.isdigit() else 0
    tot = sum(num(r[iex]) for r in data) or 1
    tots = sum(num(r[ismp]) for r in data) or 1
    print("warp instructions", tot, "samples", tots)
    seg, cur = [], None
    for n, r in enumerate(data):
        ex, sm = num(r[iex]), num(r[ismp])
        if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
            cur = {"start": n, "ex": ex, "n": 0, "inst": 0, "smp": 0, "thr": 0.0}
            seg.append(cur)
        cur["n"] += 1
        cur["inst"] += ex
        cur["smp"] += sm
        cur["thr"] += float(r[ith] or 0)
    for s in seg:
        if 100 * s["inst"] / tot >= min_pct or 100 * s["smp"] / tots >= min_pct:
            print(f"{s['start']:5d} n={s['n']:4d} exec={s['ex']:>10d} inst%={100 * s['inst'] / tot:5.1f} "
                  f"smp%={100 * s['smp'] / tots:5.1f} thr={s['thr'] / s['n']:4.1f}  "
                  f"{data[s['start']][isrc][:56]}")
