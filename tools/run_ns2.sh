set -e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rank_small|dom_count|crowd_k|select_k|merge_k|duplicate_k|variation_k|normalize_k|xcov|trmm|epilogue|transpose' --launch-skip 1500 --launch-count 120 --csv --log-file gpurun_out/ns_c2_late.csv python tools/profile_nsga2.py ei 1000 200 120 > gpurun_out/ns_c2_ncu.log 2>&1
tail -2 gpurun_out/ns_c2_ncu.log
