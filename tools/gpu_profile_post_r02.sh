#!/bin/bash
# `ncu --set full` of the posterior kernels around the triangular products at c4 (one 16384-candidate chunk).
# usage: tools/gpu_profile_post_r02.sh <tag> [env assignments for the profiled run...]
set -u
TAG=${1:-r02}; shift
OUT=gpurun_out
env "$@" python bench.py --candidates 16384 --steps 1 --warmup 3 --cpu-seconds 0.5 --extras none > $OUT/bench_post_plain_$TAG.json 2> $OUT/bench_post_plain_$TAG.err || exit 1
env "$@" ncu --set full --clock-control none -k regex:"xcov|gradsum|epilogue_kernel" -s 6 -c 6 -o $OUT/prof_post_$TAG -f \
    python bench.py --candidates 16384 --steps 1 --warmup 3 --cpu-seconds 0.5 --extras none > $OUT/ncu_post_$TAG.log 2>&1
echo "capture rc=$?"
python tools/ncu_summary.py $OUT/prof_post_$TAG.ncu-rep $OUT/ncu_posterior_kernels_${TAG}_summary.json > $OUT/ncu_post_summary_$TAG.log 2>&1
rm -f $OUT/prof_post_$TAG.ncu-rep
