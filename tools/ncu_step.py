#!/usr/bin/env python
"""Per-launch durations of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` log of tools/profile_step.py.
usage: tools/ncu_step.py LOG.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[ik].split("(")[0].replace("ipcfp::", "").replace("void ", ""), float(r[iv].replace(",", "")) / 1000) for r in rows[1:]]
starts = [i for i, l in enumerate(L) if l[0] == "k_setup"]
step = L[starts[-1]:]
tot = 0.0
for name, us in step:
    print(f"{name[:44]:44s} {us:8.1f}")
    tot += us
print(f"TOTAL {tot:.1f} us in {len(step)} launches")
