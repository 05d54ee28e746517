"""Key metrics of every kernel instance in an ncu report (ncu --set full):  python profiles/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("== " + r[h.index("Kernel Name")][:90])
    for m in METRICS:
        if m in h:
            print(f"  {m:66s} {r[h.index(m)]} {u[h.index(m)]}")
