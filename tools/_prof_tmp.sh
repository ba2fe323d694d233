set -x
CMD="python bench.py --steps 1 --warmup 3 --no-tts --no-cpu-baseline"
$CMD > gpurun_out/r2_plain_e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v8.csv $CMD > gpurun_out/r2_ncu_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfs_index_group_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_group_global_48bit $CMD > gpurun_out/r2_ncu_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfs_split_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_split_48bit $CMD > gpurun_out/r2_ncu_g.log 2>&1
export NDP_PROFILE_MAX_UNITS=30000
CMD3="python tools/ncu_probe.py rsaFACT-64bit 4096"
$CMD3 > gpurun_out/r2_probe64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfs_index_group_kernel -s 1 -c 1 -o gpurun_out/r2_dfs_group_global_64bit $CMD3 > gpurun_out/r2_ncu_h.log 2>&1
cat gpurun_out/r2_probe64.log | tail -2
ls -la gpurun_out/r2_dfs_group_global* gpurun_out/r2_dfs_split_48bit*
