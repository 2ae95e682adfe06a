"""Normalise a trace_timeline.py dump: python tools/tl.py file [max_cycles] [layers-with-chunks,...]"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 270000
show = set(int(x) for x in sys.argv[3].split(',')) if len(sys.argv) > 3 else {0, 1, 11, 12, 13, 22, 23, 24, 25}
rows = []
for l in lines:
    m = re.match(r'\s*(\d+)\s+L(\d+)\s+(.*)', l)
    if m: rows.append((int(m.group(1)), int(m.group(2)), m.group(3)))
rows = [r for r in rows if r[0] > 1e6] or rows
t0 = min(r[0] for r in rows)
for t, c, n in rows:
    if 'chunk' in n and c not in show: continue
    print(f'{t - t0:8d} L{c:<3d} {n}')
    if t - t0 > lim: break
