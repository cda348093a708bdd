"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:60]
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"{k:60s} n={v[0]:5d} total={v[1]:10.1f}us avg={v[1]/v[0]:8.1f}us {100*v[1]/tot:5.1f}%")
