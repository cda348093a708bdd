#!/bin/bash
# launch lists of the blocked fit path (plain run first, then ncu): usage tools/run_blocked_prof.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
python tools/bench_blocked.py 500 2000 > $OUT/bench_blocked_$TAG.jsonl 2> $OUT/bench_blocked_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_blocked_$TAG.csv python tools/bench_blocked.py 500 2000 > $OUT/ncu_blocked_$TAG.log 2>&1
echo "ncu rc=$?"
