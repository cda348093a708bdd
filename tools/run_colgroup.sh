#!/bin/bash
# DRAM traffic and duration of the persistent triangular product for several L2 column-group sizes (plain run first)
OUT=gpurun_out
SMALL="python bench.py --candidates 32768 --steps 1 --warmup 1 --cpu-seconds 0.2 --extras none"
for G in 2 4 8 16 32; do
  AOED_TRMM_COLGROUP=$G $SMALL > $OUT/colgroup_plain_$G.json 2> $OUT/colgroup_plain_$G.err || { echo "plain failed $G"; continue; }
  python -c "import json;l=json.load(open('$OUT/colgroup_plain_$G.json'));print('G=$G value',l['value'],'trmm ms',l['roofline']['avg_launch_ms'],'frac',l['roofline']['frac'])"
  AOED_TRMM_COLGROUP=$G ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trmm_tma -s 4 -c 2 --csv --log-file $OUT/colgroup_ncu_$G.csv $SMALL > $OUT/colgroup_ncu_$G.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader([l for l in open('$OUT/colgroup_ncu_$G.csv') if not l.startswith('==')])]
agg={}
for r in rows:
    agg.setdefault(r['ID'],{})[r['Metric Name']]=(float(r['Metric Value'].replace(',','')),r['Metric Unit'])
for k,v in agg.items(): print('G=$G', k, {n:(x,u) for n,(x,u) in v.items()})
PY
done
