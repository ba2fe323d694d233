#!/bin/bash
# the 64-bit CNF through the program on G devices: -q 4096, then -d 20 (BASELINE configs[3] verbatim)
set -u
G=${1:-8}
python tools/gen_cnf.py --named rsaFACT-64bit /tmp/rsaFACT-64bit.dimacs
( cd /tmp && time timeout 200 $OLDPWD/ndp-5_b200/bin/NDP-5_6_7 /tmp/rsaFACT-64bit.dimacs -q 4096 -g $G ) > gpurun_out/cli64_q4096_g$G.log 2>&1
grep -E "FACT|verified|BFS time:|DFS time:|NDP time:|Engine|real" gpurun_out/cli64_q4096_g$G.log
( cd /tmp && time timeout 300 $OLDPWD/ndp-5_b200/bin/NDP-5_6_7 /tmp/rsaFACT-64bit.dimacs -d 20 -g $G ) > gpurun_out/cli64_d20_g$G.log 2>&1
grep -E "FACT|verified|BFS time:|DFS time:|NDP time:|Engine|real" gpurun_out/cli64_d20_g$G.log
