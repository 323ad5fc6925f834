#!/usr/bin/env python
"""Dynamic SASS accounting of one `ncu --set full --import-source on` capture.

  python tools/ncu_sass.py gpurun_out/prof_c2.ncu-rep [warp-tasks] [--dump]

Prints the opcode histogram by executed warp instructions (exact, from the SASS page: no double counting of
inlined source lines), per-task averages when the number of warp-tasks is given, and with --dump every SASS
instruction with its execution count and stall samples (for reading hot regions)."""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    tasks = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    dump = "--dump" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    ops, samp = Counter(), Counter()
    tot = tot_s = 0
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        toks = r[i_src].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        base = op.split(".")[0]
        n, s = int(r[i_ex]), int(r[i_s])
        ops[base] += n
        samp[base] += s
        tot += n
        tot_s += s
        if dump:
            print(f"{r[0][-5:]} {n:10d} {s:6d}  {r[i_src].strip()}")
    print(f"total warp instructions {tot}, samples {tot_s}" + (f", {tot / tasks:.1f} per warp-task" if tasks else ""))
    for op, n in ops.most_common(45):
        per = f" {n / tasks:8.1f}/task" if tasks else ""
        print(f"  {op:12s} {n:12d} {100 * n / tot:6.2f}%{per}   samples {100 * samp[op] / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main()
