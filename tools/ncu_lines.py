#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
tot, stall, byfile = collections.Counter(), collections.Counter(), collections.Counter()
fname, hdr = None, None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    key = (fname, ln, r[1].strip()[:100])
    tot[key] += int(d["Instructions Executed"] or 0)
    stall[key] += int(d["# Samples"] or 0)
    byfile[fname] += int(d["Instructions Executed"] or 0)
T, S = sum(tot.values()) or 1, sum(stall.values()) or 1
print(f"kernel {kern}: {T} warp instructions, {S} stall samples")
print("--- top lines by instructions executed (inst%  samples%)")
for k, v in tot.most_common(topn): print(f"{v / T * 100:5.1f}% {stall[k] / S * 100:5.1f}%  {k[0]}:{k[1]}  {k[2]}")
print("--- top lines by stall samples (samples%  inst%)")
for k, v in stall.most_common(topn // 2): print(f"{v / S * 100:5.1f}% {tot[k] / T * 100:5.1f}%  {k[0]}:{k[1]}  {k[2]}")
print("--- by file:", {k: f"{v / T * 100:.1f}%" for k, v in byfile.most_common()})
