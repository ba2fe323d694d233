for pw in 8 32 128; do
  NDP_TRACE=1 NDP_SPLIT_PER_WARP=$pw python bench.py --workload rsaFACT-48bit:q:4096 --steps 2 --warmup 3 --no-tts --no-cpu-baseline > gpurun_out/r2_s48_pw$pw.json 2> gpurun_out/r2_s48_pw$pw.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_s48_pw$pw.json"))
print("pw=$pw", "dfs ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "bfs ms", d["e2e"]["bfs_kernel_ms"], "nodes", d["dfs"]["nodes"], "first_sat", d["dfs"]["first_sat"], "n_sat", d["dfs"]["n_sat"])
PY
  grep -m2 "ndp_split" gpurun_out/r2_s48_pw$pw.err
done
