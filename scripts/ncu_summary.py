"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of counters the design notes quote.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv > profiles/NAME.csv
"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
idx = [(w, h.index(w)) for w in WANT if w in h]
out = csv.writer(sys.stdout)
out.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
for w, i in idx:
    out.writerow([w, units[i]] + [r[i] for r in rows[2:]])
