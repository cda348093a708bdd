#!/bin/bash
# One `ncu --set full` pass over the non-dominant kernels (xcov, gradsum, epilogue from the small bench; fit / Hessian /
# HVI / rank kernels from tools/profile_others.py).  usage: tools/gpu_profile_others.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
SMALL="python bench.py --candidates 16384 --steps 1 --warmup 1 --cpu-seconds 0.2"
$SMALL > $OUT/bench_small_others_$TAG.json 2> $OUT/bench_small_others_$TAG.err && python tools/profile_others.py > $OUT/profile_others_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k regex:"xcov|gradsum|epilogue" -s 3 -c 3 -o $OUT/prof_post_$TAG -f $SMALL > $OUT/ncu_post_$TAG.log 2>&1
echo "posterior kernels rc=$?"
ncu --set full --clock-control none -k regex:"lml_blk|hess_kernel|dkcol|hvc_kernel|rank_small|argmax" -c 8 -o $OUT/prof_misc_$TAG -f python tools/profile_others.py > $OUT/ncu_misc_$TAG.log 2>&1
echo "misc kernels rc=$?"
