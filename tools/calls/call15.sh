cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_tma.sum,sm__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for st in barrier branch_resolving dispatch_stall drain lg_throttle long_scoreboard math_pipe_throttle membar mio_throttle misc no_instruction not_selected selected short_scoreboard sleeping tex_throttle wait; do M=$M,smsp__warp_issue_stalled_${st}_per_warp_active.pct; done
timeout 1700 ncu --replay-mode application --clock-control none -k regex:executor_kernel -c 2 --metrics $M --csv --log-file gpurun_out/r02_ncu_executor_lap3d_100.csv python tools/ncu_factor.py lap3d_100 2>&1 | tail -3
wc -l gpurun_out/r02_ncu_executor_lap3d_100.csv
