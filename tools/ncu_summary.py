"""summarise a .ncu-rep (raw page) into the handful of counters DESIGN.md argues with"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__sectors_read.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:150])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:85s} {r[i]:>20s} {units[i]}")
