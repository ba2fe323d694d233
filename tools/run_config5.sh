#!/bin/bash
# BASELINE configs[4] (index 5 in SURVEY §8d): UNSAT prime-input CNF, exhaustive DFS, straight vs -s / -r.
set -u
G=${1:-1}
W=/tmp/cfg5; rm -rf $W; mkdir -p $W; cd $W
python $OLDPWD/tools/gen_cnf.py --named prime-32bit $W/prime-32bit.dimacs
BIN=$OLDPWD/ndp-5_b200/bin/NDP-5_6_7
( time $BIN prime-32bit.dimacs -q 1024 -x -g $G ) > straight.log 2>&1; echo "straight rc=$?"
( time $BIN prime-32bit.dimacs -q 1024 -x -g $G -s ) > save.log 2>&1; echo "save rc=$?"
J=$(ls bfs_*prime-32bit*.json | head -1); ls -la $J
( time $BIN prime-32bit.dimacs -q 1024 -x -g $G -r $J ) > resume.log 2>&1; echo "resume rc=$?"
for f in straight save resume; do echo "== $f"; grep -E "Prime|FACT|BFS time:|DFS time:|Queue Size|Engine|real" $f.log; done
cp straight.log $OLDPWD/gpurun_out/cfg5_straight_g$G.log; cp resume.log $OLDPWD/gpurun_out/cfg5_resume_g$G.log; cp save.log $OLDPWD/gpurun_out/cfg5_save_g$G.log
