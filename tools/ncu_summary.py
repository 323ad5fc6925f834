#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small text file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof_c2.ncu-rep [megapixels-per-launch] > profiles/r01_c2_fused.txt

Prints, per captured kernel launch: duration, DRAM bytes, instruction counts (total and per pixel), pipe
utilisation, issue-slot utilisation, shared-memory wavefronts, the opcode histogram and the hottest source lines
(needs -lineinfo).  Runs here on the CPU box (ncu -i only reads the report)."""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    mp = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        print(f"== {name}")
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                vals[h] = row[i]
                print(f"  {h:70s} {row[i]:>16s} {units[i]}")
        if mp:
            try:
                wi = float(vals["smsp__inst_executed.sum"])
                us = float(vals["gpu__time_duration.sum"])
                print(f"  -> {wi / (mp * 1e6):.2f} warp-instructions/px = {32 * wi / (mp * 1e6):.0f} thread-instruction slots/px; "
                      f"{mp / us:.3f} MP/us")
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
                rd = float(vals["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
                wr = float(vals["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
                print(f"  -> dram traffic {rd + wr:.2f} MB per launch vs algorithmic bytes (see DESIGN.md)")
            except Exception as e:  # noqa: BLE001
                print("  (derived figures unavailable:", e, ")")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, lines, ops, tot_i, tot_s = None, [], Counter(), 0, 0
    opsamp = Counter()
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 7 and r[0].isdigit() and r[6].isdigit():
            lines.append((cur, int(r[0]), r[1].strip(), int(r[6]), int(r[7])))
        elif len(r) > 7 and r[0] == "" and r[2].startswith("0x") and r[7].isdigit():
            toks = r[3].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.split(".")[0]
            ops[op] += int(r[7])
            opsamp[op] += int(r[6]) if r[6].isdigit() else 0
            tot_i += int(r[7])
            tot_s += int(r[6]) if r[6].isdigit() else 0
    if tot_i:
        print("\n-- opcode histogram (warp instructions executed, first captured kernel's source page)")
        for op, n in ops.most_common(28):
            print(f"  {op:10s} {n:12d} {100 * n / tot_i:6.2f}%   stall samples {100 * opsamp[op] / max(tot_s, 1):5.1f}%")
    if lines:
        ts = sum(x[3] for x in lines) or 1
        ti = sum(x[4] for x in lines) or 1
        print("\n-- hottest source lines (share of stall samples | share of instructions)")
        for x in sorted(lines, key=lambda o: -o[3])[:30]:
            print(f"  {x[0]}:{x[1]:<4d} {100 * x[3] / ts:5.1f}% | {100 * x[4] / ti:5.1f}%   {x[2][:100]}")


if __name__ == "__main__":
    main()
