#!/bin/bash
# 8-GPU evidence: default bench, BASELINE configs[2] (48-bit -q 65536, exhaustive), and the 64-bit CNF through the CLI.
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 default rc=$?"; tail -c 600 gpurun_out/bench_n8.json
$TR bench.py --gpus 8 --steps 3 --warmup 3 --workload rsaFACT-48bit:q:65536 > gpurun_out/bench48_n8.json 2> gpurun_out/bench48_n8.err; echo "n8 48bit rc=$?"; tail -c 900 gpurun_out/bench48_n8.json
python tools/gen_cnf.py --named rsaFACT-64bit /tmp/rsaFACT-64bit.dimacs
( cd /tmp && time timeout 300 $OLDPWD/ndp-5_b200/bin/NDP-5_6_7 /tmp/rsaFACT-64bit.dimacs -q 4096 -g 8 ) > gpurun_out/cli64_g8.log 2>&1; echo "cli64 rc=$?"; tail -45 gpurun_out/cli64_g8.log
