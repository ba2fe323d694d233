set -x
CMD2="python bench.py --workload rsaFACT-32bit:q:16384 --steps 1 --warmup 3 --no-tts --no-cpu-baseline"
NDP_DFS_GROUPS=4 $CMD2 > gpurun_out/r2_plain_g4.log 2>&1 && \
NDP_DFS_GROUPS=4 ncu --set full --clock-control none --import-source on -k regex:dfs_index_group_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_group8_32bit $CMD2 > gpurun_out/r2_ncu_g4.log 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-tts --no-cpu-baseline"
NDP_DFS_GROUPS=2 $CMD > gpurun_out/r2_plain_g2.log 2>&1 && \
NDP_DFS_GROUPS=2 ncu --set full --clock-control none --import-source on -k regex:dfs_index_group_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_group16_48bit $CMD > gpurun_out/r2_ncu_g2.log 2>&1
ls -la gpurun_out/r2_dfs_group*
