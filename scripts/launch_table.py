"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and mean time, share."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ki])
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[ui], 1.0)
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
total = sum(tot.values())
print('| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('| `%s` | %d | %.0f | %.1f | %.1f %% |' % (k, cnt[k], v, v / cnt[k], 100 * v / total))
print('total %.1f ms over %d launches' % (total / 1e3, sum(cnt.values())))
