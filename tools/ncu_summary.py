"""Summarise one kernel of an .ncu-rep (ncu --set full) into the CSV format kept under profiles/.
usage: python tools/ncu_summary.py <file.ncu-rep> "<label>" > profiles/rNN_ncu_<kernel>_summary.csv"""
import csv
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "sm__icc_request_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
label = sys.argv[2] if len(sys.argv) > 2 else vals[names.index("Kernel Name")]
print("kernel,metric,unit,value")
for n, u, v in zip(names, units, vals):
    if n in WANT:
        print(f"{label},{n},{u},{v}")
