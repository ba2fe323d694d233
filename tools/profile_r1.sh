#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + --set full captures of the hot kernels.
# Each ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-tts --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v6.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
C1="python tools/dfs_reps.py rsaFACT-32bit:q:16384 2"
$C1 > gpurun_out/plain_c1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dfs_index_kernel -s 1 -c 1 -f -o gpurun_out/r1_dfs_index_v6 $C1 > gpurun_out/ncu_c1.log 2>&1
echo "dfs full rc=$?"
$C1 > gpurun_out/plain_c1b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bfs_index_persistent -c 1 -f -o gpurun_out/r1_bfs_persistent_v6 $C1 > gpurun_out/ncu_c1b.log 2>&1
echo "bfs full rc=$?"
C2="python tools/dfs_reps.py rsaFACT-32bit:d:12 2"
$C2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dfs_split_kernel -s 1 -c 1 -f -o gpurun_out/r1_dfs_split_v6 $C2 > gpurun_out/ncu_c2.log 2>&1
echo "split full rc=$?"
ls -la gpurun_out/*.ncu-rep
