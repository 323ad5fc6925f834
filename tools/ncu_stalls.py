#!/usr/bin/env python
"""Stall reasons of one `ncu --set full --import-source on` capture, summed over all SASS instructions and per class of
execution count (instructions executed equally often belong to the same loop / function):
  python tools/ncu_stalls.py gpurun_out/prof_spec_c2.ncu-rep [min-share-percent]"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None
tot = Counter()
per = defaultdict(Counter)
ninst = Counter()
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        hdr = r
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i_ex, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    ex = int(r[i_ex])
    ninst[ex] += 1
    for i in cols:
        v = int(r[i]) if r[i].isdigit() else 0
        tot[hdr[i]] += v
        per[ex][hdr[i]] += v
allsamp = sum(tot.values()) or 1
print("all instructions:", ", ".join(f"{k[6:]} {100 * v / allsamp:.1f}%" for k, v in tot.most_common(8)))
for ex, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    share = 100 * sum(c.values()) / allsamp
    if share < min_share:
        continue
    print(f"executed {ex:>8d} x ({ninst[ex]:4d} instructions): {share:5.1f}% of samples:",
          ", ".join(f"{k[6:]} {100 * v / allsamp:.1f}" for k, v in c.most_common(5)))
