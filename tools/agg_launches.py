"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:100]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v; n += 1
tot = sum(v[1] for v in agg.values())
print('launches %d, total %.1f us' % (n, tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%10.1f us %5.1f%% n=%4d avg %8.1f  %s' % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k))
