#!/bin/bash
# BASELINE configs[3] verbatim: 64-bit CNF, -d 20 (ONE subproblem), 8 x B200, early exit.  Bounded by a timeout.
set -u
G=${1:-8}; T=${2:-280}
python tools/gen_cnf.py --named rsaFACT-64bit /tmp/rsaFACT-64bit.dimacs
( cd /tmp && time NDP_TRACE=1 timeout $T $OLDPWD/ndp-5_b200/bin/NDP-5_6_7 /tmp/rsaFACT-64bit.dimacs -d 20 -g $G ) > gpurun_out/cli64_d20_g$G.log 2>&1
echo "rc=$?"; grep -vE "persistent" gpurun_out/cli64_d20_g$G.log | tail -45
