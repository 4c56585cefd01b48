"""Print the key metrics of every kernel in an .ncu-rep (uses `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.avg.per_cycle_active',
 'smsp__issue_active.avg.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'sm__sass_thread_inst_executed_op_fp32_pred_on.sum','smsp__sass_thread_inst_executed_op_ffma_pred_on.sum','smsp__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__sass_thread_inst_executed_op_fadd_pred_on.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:60], 'id', r[0])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w); print(f'   {w:88s} {r[i]:>18s} {units[i]}')
