#!/bin/bash
# usage: tools/run_sanitizer.sh <racecheck|memcheck> <tag>   (plain run first; then under compute-sanitizer)
TOOL=${1:-memcheck}; TAG=${2:-r02}; OUT=gpurun_out
python tools/sanitize_small.py > $OUT/sanitize_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/sanitize_plain_$TAG.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 30 python tools/sanitize_small.py > $OUT/sanitizer_${TOOL}_$TAG.log 2>&1
echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" $OUT/sanitizer_${TOOL}_$TAG.log | tail -12
