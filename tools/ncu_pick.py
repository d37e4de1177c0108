"""Print selected metrics of an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv), one block per launch."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:110], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(h.endswith(w) for w in WANT) or any(e in h for e in extra):
            print(f"   {h:90s} {r[i]} {units[i]}")
