set -x
CMD="python bench.py --steps 1 --warmup 3 --no-tts --no-cpu-baseline"
$CMD > gpurun_out/r2_plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v7.csv $CMD > gpurun_out/r2_ncu_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfs_index_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_index_v7_48bit $CMD > gpurun_out/r2_ncu_b.log 2>&1
CMD2="python bench.py --workload rsaFACT-32bit:q:16384 --steps 1 --warmup 3 --no-tts --no-cpu-baseline"
$CMD2 > gpurun_out/r2_plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfs_index_kernel -s 3 -c 1 -o gpurun_out/r2_dfs_index_v7_32bit $CMD2 > gpurun_out/r2_ncu_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bfs_index_persistent_kernel -s 3 -c 1 -o gpurun_out/r2_bfs_persistent_v7_32bit $CMD2 > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
