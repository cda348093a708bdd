#!/bin/bash
# Round profile recipe (B200_PROFILING.md): plain run first, then the launch list and one full capture of the top kernel.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
SMALL="python bench.py --candidates 32768 --steps 1 --warmup 1 --cpu-seconds 0.2"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json
$SMALL > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"xcov|trmm|gradsum|epilogue|transpose_x|hess|dkcol|gram" -c 120 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$SMALL > $OUT/bench_small2_$TAG.json 2>> $OUT/bench_small_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:trmm -s 4 -c 2 -o $OUT/prof_trmm_$TAG -f $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la $OUT | tail -12
