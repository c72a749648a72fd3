"""DRAM bytes per launch of a kernel from an `ncu --set full` report -> one entry of profiles/traffic.json, the file
bench.py reads `roofline.traffic` from (key: "<workload>:lps<local LPs>:world<N>"; a run without a matching capture
prints null).

usage: python tools/ncu_traffic.py report.ncu-rep <workload> <local LPs> <world> [kernel regex]
The capture must be of the SAME workload and partition size as the bench line it annotates (a shortened stop time is fine:
the per-launch traffic of a steady-state super-step does not depend on how many steps follow)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, workload, n_local, world = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
pat = re.compile(sys.argv[5] if len(sys.argv) > 5 else "k_step|k_run_groups")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = wr = dur = 0.0
n = 0
for r in rows[2:]:
    if not pat.search(r[col["Kernel Name"]]):
        continue
    rd += float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr += float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    dur += float(r[col["gpu__time_duration.sum"]])
    n += 1
if not n:
    sys.exit("no matching kernel in the report")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "traffic.json")
try:
    table = json.load(open(path))
except Exception:
    table = {}
key = f"{workload}:lps{n_local}:world{world}"
table[key] = {"dram_bytes_per_launch": (rd + wr) / n, "dram_read": rd / n, "dram_write": wr / n, "launches_captured": n,
              "duration_under_ncu": dur / n, "duration_unit": units[col["gpu__time_duration.sum"]],
              "source": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(key, table[key])
