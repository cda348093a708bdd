#!/bin/bash
# One `ncu --set full` pass over the kernels of the device-resident NSGA-II loop at steady state (pop 1000, EI).
# usage: tools/gpu_profile_nsga2.sh <tag>     (run only after `python tools/profile_nsga2.py ...` exited 0 without ncu)
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python tools/profile_nsga2.py ei 1000 200 120"
AOED_NSGA2_GRAPH=0 $CMD > $OUT/nsga2_plain_$TAG.log 2>&1 &&
AOED_NSGA2_GRAPH=0 ncu --set full --clock-control none --import-source on \
  -k regex:"small_eval|dom_count|rank_small|crowd_pos|select_k|duplicate_k|variation_k|merge_k" --launch-skip 800 --launch-count 8 \
  -o $OUT/prof_nsga2_$TAG -f $CMD > $OUT/ncu_nsga2_$TAG.log 2>&1
echo "nsga2 kernels rc=$?"
