"""Print CTA 0's MMA-issuer / epilogue timeline for the first ray group (developer tool).
events: 0 mma layer start | 8 mma layer issued (+ cycles the issuer was blocked on operands / on the weight ring) | 10 epilogue: accumulator ready
        11+j epilogue chunk j stored+signalled | 15 epilogue (head layer) done"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smpl_nerf_b200 import engine
wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
precision = 1 if (len(sys.argv) > 2 and sys.argv[2] == 'fast') else 0
w = bench.WORKLOADS[wl]
coarse, fine, warp, pe, de, he = bench.build_models(w)
dev = torch.device('cuda:0')
coarse, fine = coarse.to(dev), fine.to(dev)
warp = warp.to(dev) if warp is not None else None
args = bench.make_args(w)
data = [t.to(dev) for t in bench.make_views(w, 0, 1)[0]]
for _ in range(2):
  with torch.no_grad():
    out = engine.render(w['kind'], coarse, fine, warp, args, pe, de, he, data, precision=precision, trace_cap=6000)
torch.cuda.synchronize()
tr = out['trace'].cpu()
n = min(int(tr[1]), int(tr[0]))
ev = tr[2:2 + 3 * n].view(n, 3).tolist()
waits = {(e, c): t for e, c, t in ev if e in (20, 21)}
ev = [e for e in ev if e[0] not in (20, 21)]
ev.sort(key=lambda e: e[2])
t0 = ev[0][2]
names = {0: 'M start', 8: 'M issued', 10: 'E acc-ready', 15: 'E head done', 30: 'T tile top', 31: 'T points loaded', 32: 'T warp-PE written',
         33: 'T warp-PE published', 35: 'T warp heads exchanged', 36: 'T warped pos + dirs done', 37: 'T PE written', 38: 'T PE published',
         39: 'T heads exchanged (tile end)', 40: 'R pass raw complete', 41: 'R ray stage done'}
last = {}
nl = 12 * 4 + 2
for e, ctr, t in ev:
    if ctr >= nl: continue
    nm = names.get(e, f'M ops{e-1} rdy' if 1 <= e <= 7 else f'E chunk{e-11}')
    extra = f'   (issuer blocked: operands {waits.get((20, ctr), 0)}, weight ring {waits.get((21, ctr), 0)} cycles)' if e == 8 else ''
    print(f'{t - t0:9d}  L{ctr:<3d} {nm}{extra}')
