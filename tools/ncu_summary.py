#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {v / 1e3:12.1f} {100 * v / total:6.1f}%")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {total / 1e3:12.1f}")
