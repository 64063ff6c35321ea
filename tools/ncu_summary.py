"""Key metrics of every kernel in an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
print("# %s  (ncu --set full --clock-control none: serialised, cold cache -- durations are NOT bench numbers)" % sys.argv[1])
for r in data:
    print("\n== %s" % r[ik][:150])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-95s %14s %s" % (k, r[i], units[i]))
