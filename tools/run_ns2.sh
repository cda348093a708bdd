mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
AOED_NSGA2_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rank_small|dom_count|crowd_pos|select_k|merge_k|duplicate_k|variation_k|normalize_k|xcov|trmm|epilogue|transpose' --launch-skip 1500 --launch-count 330 --csv --log-file gpurun_out/launches_nsga2_r01.csv python tools/profile_nsga2.py ei 1000 200 120 > gpurun_out/ns_c2_ncu.log 2>&1
tail -1 gpurun_out/ns_c2_ncu.log
