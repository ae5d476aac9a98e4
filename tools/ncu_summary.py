import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__warps_eligible.avg.per_cycle_active','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum']
idx={h:i for i,h in enumerate(hdr)}
seen=set()
filt = sys.argv[2] if len(sys.argv)>2 else ''
for r in rows[2:]:
    name=r[idx['Kernel Name']][:40]
    if name in seen or filt not in name: continue
    seen.add(name)
    print('==',name)
    for w in want:
        if w in idx: print(f"   {w:90s} {r[idx[w]]:>18s} {units[idx[w]]}")
