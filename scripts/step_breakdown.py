"""One full step out of an ncu launch list (delimited by consecutive topk_kernel launches): per-kernel share + ordered list."""
import csv, re, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
seq = []
for i, r in enumerate(rows):
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    seq.append((i, name[:70], r.get('Grid Size'), r.get('Block Size'), v))
idx = [i for i, n, _, _, _ in seq if n.startswith('topk_kernel')]
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(idx) - 2
a, b = idx[which], idx[which + 1]
step = seq[a:b]
tot = sum(s[4] for s in step)
print('%d launches in the step, %.1f us of kernel time (cold-cache, serialised)' % (len(step), tot))
agg = collections.defaultdict(lambda: [0, 0.0])
for s in step:
    agg[s[1]][0] += 1; agg[s[1]][1] += s[4]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print('%-72s n=%3d %9.1f us %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
if len(sys.argv) > 4:
    for s in step:
        print('%4d %-60s %-14s %-12s %8.1f' % s)
