for pw in 1 2 4 16; do
  NDP_TRACE=1 NDP_SPLIT_PER_WARP=$pw python bench.py --steps 2 --warmup 3 --no-tts --no-cpu-baseline > gpurun_out/r2_pw.json 2> gpurun_out/r2_pw.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_pw.json'))
print('per_warp $pw dfs ms', d['e2e']['dfs_kernel_ms'])
"
  grep -m1 "ndp_split" gpurun_out/r2_pw.err
  grep "ndp_dfs\]" gpurun_out/r2_pw.err | tail -1
done
