#!/usr/bin/env python3
"""Print the key metrics of every kernel in an `ncu --page raw --csv` dump (one column per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
idx = {h: i for i, h in enumerate(hdr)}
for w in want:
    if w in idx:
        print(f"{w[:70]:70s} [{units[idx[w]][:8]:>8s}] " + " | ".join(f"{r[idx[w]][:14]:>14s}" for r in rows[2:]))
