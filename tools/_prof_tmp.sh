export NDP_PROFILE_MAX_STEPS=20000
CMD3="python tools/ncu_probe.py rsaFACT-64bit 4096"
timeout 150 $CMD3 > gpurun_out/r2_probe64.log 2>&1 && \
timeout 400 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:dfs_index_group_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2_dfs_group_global_64bit_metrics.csv $CMD3 > gpurun_out/r2_ncu_h.log 2>&1
tail -1 gpurun_out/r2_probe64.log; tail -1 gpurun_out/r2_ncu_h.log; tail -16 gpurun_out/r2_dfs_group_global_64bit_metrics.csv | cut -d, -f12-
