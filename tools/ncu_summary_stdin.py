"""summarise `ncu -i x.ncu-rep --page raw --csv` (stdin) into the counters DESIGN.md argues with"""
import csv, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:120])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:85s} {r[i]:>20s} {units[i]}")
