"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = row['Kernel Name']
    k = k.split('(')[0][:66]
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share of GPU time |")
print("|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
print("\ntotal GPU time in listed launches: %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
