#!/bin/bash
# `ncu --set full` of the kernels of the blocked fit path and of the HVI loop (plain runs first).  usage: tools/gpu_profile_fit_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
python tools/bench_blocked.py 2000 > $OUT/bench_blocked_plain_$TAG.jsonl 2> $OUT/bench_blocked_plain_$TAG.err || exit 1
# 48-problem gradient batch at N=2000: skip the 3-problem batches, capture a window over the Cholesky steps / inverse / K^-1 / gradient
ncu --set full --clock-control none -k regex:"tgemm|potrf_diag|kmat_kernel|lml_grad_kernel" -s 214 -c 30 \
    -o $OUT/prof_blockedfit_$TAG -f python tools/bench_blocked.py 2000 > $OUT/ncu_blockedfit_$TAG.log 2>&1
echo "blocked fit capture rc=$?"
python tools/bench_hvi.py > $OUT/bench_hvi_plain_$TAG.jsonl 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"hvi_round|dominated_by|hvc_kernel" -c 6 -o $OUT/prof_hvi_$TAG -f python tools/bench_hvi.py > $OUT/ncu_hvi_$TAG.log 2>&1
echo "hvi capture rc=$?"
# the reports are too large to travel (2 MB per kernel): summarise on the box, keep only the JSON
python tools/ncu_summary.py $OUT/prof_blockedfit_$TAG.ncu-rep $OUT/ncu_blocked_fit_kernels_${TAG}_summary.json > $OUT/ncu_blockedfit_summary_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_hvi_$TAG.ncu-rep $OUT/ncu_hvi_kernels_${TAG}_summary.json > $OUT/ncu_hvi_summary_$TAG.log 2>&1
rm -f $OUT/prof_blockedfit_$TAG.ncu-rep $OUT/prof_hvi_$TAG.ncu-rep
ls -la $OUT | tail -8
