#!/usr/bin/env python3
"""Markdown summary of one kernel of an .ncu-rep (ncu --set full): duration, instruction and issue statistics, stall
reasons, DRAM traffic.  Usage: python profiles/ncu_summary.py <report.ncu-rep> [kernel-substring]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALLS = ["barrier", "wait", "short_scoreboard", "long_scoreboard", "no_instruction", "not_selected", "branch_resolving",
          "math_pipe_throttle", "mio_throttle", "lg_throttle", "dispatch_stall", "membar", "sleeping", "drain"]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        if sub in r[name_col]:
            break
    else:
        raise SystemExit("kernel not found")
    col = {h: i for i, h in enumerate(hdr)}
    print("kernel: `%s`\n" % r[name_col][:140])
    print("| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in col:
            print("| %s | %s %s |" % (label, r[col[key]], units[col[key]]))
    print("\n| stall reason (warp-cycles per issued instruction) | value |\n|---|---|")
    for s in STALLS:
        key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if key in col:
            print("| %s | %.2f |" % (s, float(r[col[key]])))


if __name__ == "__main__":
    main()
