#!/bin/bash
# The north-star target: the FULL subproblem frontier of the 64-bit CNF (-q 4096) resolved on G devices, no early exit.
set -u
G=${1:-8}; T=${2:-110}
python tools/gen_cnf.py --named rsaFACT-64bit /tmp/rsaFACT-64bit.dimacs
( cd /tmp && time timeout $T $OLDPWD/ndp-5_b200/bin/NDP-5_6_7 /tmp/rsaFACT-64bit.dimacs -q 4096 -g $G -x ) > gpurun_out/cli64_q4096_exhaustive_g$G.log 2>&1
echo "rc=$?"; grep -E "FACT|verified|BFS time:|DFS time:|NDP time:|Queue Size|Engine|real" gpurun_out/cli64_q4096_exhaustive_g$G.log
