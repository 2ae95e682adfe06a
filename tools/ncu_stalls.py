"""Per-instruction warp-stall summary of one kernel from an ncu report (source page).  Usage: python tools/ncu_stalls.py <report.ncu-rep> [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
print(rows[start - 1][:2])
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top_n]:
    st = sorted(((s, int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda kv: -kv[1])[:3]
    print('%5d %6s %8s  %-70s %s' % (data.index(r), r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:70], st))
agg = {}
for r in data:
    for s in stalls:
        agg[s] = agg.get(s, 0) + int(r[ix[s]])
print(sorted(agg.items(), key=lambda kv: -kv[1]))
