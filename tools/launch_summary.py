"""Per-kernel totals of ONE training step from an ncu launch list with DRAM bytes (tools/gpu_round2.sh: launches_train.csv).
Usage: python tools/launch_summary.py profiles/r2/launches_train.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
L = OrderedDict()
for r in rows:
    d = L.setdefault(int(r[0]), {'name': r[4].split('(')[0].replace('void ', '')})
    d[r[12]] = float(r[14].replace(',', ''))
    d[r[12] + '_u'] = r[13]
ks = list(L.values())
names = [k['name'] for k in ks]
period = next(p for p in range(20, len(ks)) if all(names[i] == names[i + p] for i in range(len(ks) - p)))
unit = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot, by, cnt = {}, {}, {}
for k in ks[:period]:
    n = k['name']
    tot[n] = tot.get(n, 0) + k['gpu__time_duration.sum'] / 1e3
    by[n] = by.get(n, 0) + sum(k[m] * unit[k[m + '_u']] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    cnt[n] = cnt.get(n, 0) + 1
print('%d launches per step; kernel time %.1f us; DRAM traffic %.2f GB (cold caches: ncu serialises and flushes between launches)'
      % (period, sum(tot.values()), sum(by.values()) / 1e9))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('%-34s x%-3d %8.1f us %7.2f GB %5.2f TB/s' % (n, cnt[n], t, by[n] / 1e9, by[n] / t / 1e6))
