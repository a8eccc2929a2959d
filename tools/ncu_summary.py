"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [substr ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers", "launch__occupancy",
                        "sm__warps_active.avg.pct", "inst_executed_pipe", "issue_active", "smsp__inst_executed.sum", "warp_issue_stalled",
                        "dram__throughput", "sm__throughput.avg.pct", "cycles_elapsed.max", "bank_conflicts", "smsp__warps_eligible",
                        "pipe_fma", "pipe_alu", "pipe_xu", "sm__cycles_active.avg", "clock_rate", "achieved_occupancy", "thread_inst_executed"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print(f"  {h:100s} {units[i]:14s} {r[i]}")
