"""Print the handful of ncu raw-page metrics we steer by. Usage: ncu_keys.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
d = dict(zip(rows[0], rows[2]))
print(d.get("Kernel Name"), d.get("Block Size"), d.get("Grid Size"))
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_warps", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "lts__t_sector_hit_rate.pct"]:
    print("  %-90s %s" % (k, d.get(k)))
for k in rows[0]:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(d[k]) > 0.2:
        print("  %-90s %s" % (k, d[k]))
