import csv,sys,subprocess
out=subprocess.run(["ncu","-i",sys.argv[1],"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','launch__grid_size','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__cycles_active.avg']
for r in rows[2:]:
    for w in want:
        if w in hdr: print(w, '=', r[hdr.index(w)])
    print('---')
