"""ncu report -> the compact per-kernel summary CSV kept under profiles/:
    python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/rNN_kernels_ncu_full_summary.csv"""
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed_op_branch.sum", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head = rows[0]
stalls = [k for k in head if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
          and "not_issued" not in k]
cols = [k for k in KEEP if k in head] + stalls
idx = [head.index(k) for k in cols]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
print("wrote", sys.argv[2], len(rows) - 2, "kernels")
