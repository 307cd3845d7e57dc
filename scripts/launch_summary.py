"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    rows.append(row)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    name = re.sub(r'^void ', '', row['Kernel Name'])
    name = re.sub(r'\(.*', '', name)
    name = name.replace('<unnamed>::', '')[:78]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('%d launches, total %.1f us' % (len(rows), tot))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-80s n=%4d %10.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
