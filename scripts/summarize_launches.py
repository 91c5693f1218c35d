"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
data = [r for r in rows[hi + 1:] if len(r) > mv and r[mn].startswith('gpu__time_duration')]
agg = collections.OrderedDict()
for r in data:
    name = r[kn].split('(')[0].replace('void ', '').replace('isg::', '')
    t = float(r[mv].replace(',', '')) / 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f'{len(data)} launches, {tot / 1e3:.2f} ms total')
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:10.1f} us {100 * t / tot:5.1f}%  x{n:<4d} {name[:80]}')
