for cfg in "4 32" "4 48"; do set -- $cfg
  NDP_DFS_GROUPS=$1 NDP_DFS_GLOBAL_STATE=1 NDP_INDEX_WARPS_PER_SM=$2 python bench.py --steps 2 --warmup 3 --no-tts --no-cpu-baseline > gpurun_out/r2_gs.json 2> gpurun_out/r2_gs.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_gs.json'))
print('48bit global state G=$1 warps/SM $2 dfs ms', d['e2e']['dfs_kernel_ms'])
"
  NDP_DFS_GROUPS=$1 NDP_DFS_GLOBAL_STATE=1 NDP_INDEX_WARPS_PER_SM=$2 python bench.py --workload rsaFACT-32bit:q:16384 --steps 5 --warmup 3 --no-tts --no-cpu-baseline > gpurun_out/r2_gs.json 2> gpurun_out/r2_gs.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_gs.json'))
print('32bit global state G=$1 warps/SM $2 dfs ms', d['e2e']['dfs_kernel_ms'])
"
done
