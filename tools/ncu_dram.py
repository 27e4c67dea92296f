"""DRAM traffic per launch of a kernel from an `ncu --set full` report:
python tools/ncu_dram.py REPORT KERNEL_REGEX OUT.json"""
import csv
import json
import subprocess
import sys

rep, pattern, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + pattern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]


def column(name):
    i = head.index(name)
    scale = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in data]


rd, wr = column("dram__bytes_read.sum"), column("dram__bytes_write.sum")
t = [float(r[head.index("gpu__time_duration.sum")].replace(",", "")) for r in data]
doc = {"kernel": pattern, "launches": len(data),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(data),
       "dram_read_per_launch": sum(rd) / len(data), "dram_write_per_launch": sum(wr) / len(data),
       "duration_unit": units[head.index("gpu__time_duration.sum")], "durations": t,
       "source": rep}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc))
