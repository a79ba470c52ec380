"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    agg[row['Kernel Name'][:70]].append(float(row['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'n':>4s} {'avg_us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} {len(v):4d} {sum(v)/len(v)/1000:9.1f} {sum(v)/tot*100:6.1f}%")
