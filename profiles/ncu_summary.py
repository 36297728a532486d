"""Trim `ncu --page raw --csv` exports to the columns the roofline argument uses.   python profiles/ncu_summary.py out.csv in1.csv [in2.csv ...]"""
import csv
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__cycles_active.avg"]
out = csv.writer(open(sys.argv[1], "w", newline=""))
first = True
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    if first:
        out.writerow([hdr[i] for i in idx])
        out.writerow([units[i] for i in idx])
        first = False
    for r in data:
        out.writerow([r[i] for i in idx])
