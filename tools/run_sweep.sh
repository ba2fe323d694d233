#!/bin/bash
# one bench line per workload at N GPUs (N = $1): default (32-bit SAT), UNSAT prime (configs 5), 48-bit (configs 3)
set -u
N=$1
if [ "$N" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N"; fi
$L bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_32bit_q16384_n$N.json 2> gpurun_out/sweep_a_n$N.err; echo "rc=$?"
$L bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload prime-32bit:q:16384 > gpurun_out/sweep_prime32_q16384_n$N.json 2> gpurun_out/sweep_b_n$N.err; echo "rc=$?"
$L bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --workload rsaFACT-48bit:q:65536 > gpurun_out/sweep_48bit_q65536_n$N.json 2> gpurun_out/sweep_c_n$N.err; echo "rc=$?"
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/sweep_*_n%s.json" % os.environ.get("NN","*"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "n_gpus", d["n_gpus"], "dfs ms/step %.3f"%d["ms_per_step"], "value %.0f"%d["value"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], "n_sat", d["dfs"]["n_sat"])
    except Exception as e:
        print(f, "unreadable", e)
PY
