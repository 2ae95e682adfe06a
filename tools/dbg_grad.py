import sys, copy, torch
sys.path.insert(0, '/root/repo')
from tests import helpers as H
from tests import test_gpu_train as T
from oracle import nerf_oracle as O
from smpl_nerf_b200 import engine
kind = sys.argv[1]
loop_cfg = len(sys.argv) > 2 and sys.argv[2] == 'loop'      # the configuration of test_solver_loop_tracks_the_reference_loop, free-running sampler
if loop_cfg:
    from smpl_nerf_b200 import scene
    nets = T._build(kind, 21, 4, (2,), 'dense')
    args = O.make_args(number_fine_samples=32, sigma_noise_std=0.)
    rays = scene.make_rays(24, 24, 32, seed=5, with_colours=True, arm_angle_deg=35.0)
    data = scene.data_list(rays, kind, torch.arange(0, 576, 4))
else:
    nets = T._build(kind, 7, 8, (4,), 'dense')
    args = O.make_args(number_fine_samples=64)
    data = T._rays(kind, 8, 8, 32, 11)
with torch.no_grad():
    z_all = None if loop_cfg else H.run_oracle(kind, nets, args, data)['z_all']
n64 = [copy.deepcopy(m).double() if m is not None else None for m in nets[:3]]
d64 = [t.double() for t in data]
o64 = H.run_oracle(kind, (n64[0], n64[1], n64[2]) + tuple(nets[3:]), args, d64, z_all_in=z_all)
T._loss((o64['rgb'], o64['rgb_fine']), d64[-1]).backward()
n32 = [copy.deepcopy(m) if m is not None else None for m in nets[:3]]
o32 = H.run_oracle(kind, (n32[0], n32[1], n32[2]) + tuple(nets[3:]), args, data, z_all_in=z_all)
T._loss((o32['rgb'], o32['rgb_fine']), data[-1]).backward()
gnets, gdata = H.to_cuda(nets, data)
for m in gnets[:3]:
    if m is not None: m.train()
out = engine.render(kind, gnets[0], gnets[1], gnets[2], args, gnets[3], gnets[4], gnets[5], gdata, z_all_in=None if z_all is None else z_all.to('cuda:0'))
T._loss((out['rgb'], out['rgb_fine']), gdata[-1]).backward()
for name, net, ref, r32 in zip(('coarse', 'fine', 'warp'), gnets[:3], n64, n32):
    if net is None: continue
    for (pn, p), (_, q), (_, q32) in zip(net.named_parameters(), ref.named_parameters(), r32.named_parameters()):
        g, w = p.grad.double().cpu(), q.grad
        rel = float((g - w).norm() / (w.norm() + 1e-30)); floor = float((q32.grad.double() - w).norm() / (w.norm() + 1e-30))
        sign = float((torch.sign(g) != torch.sign(q32.grad.double())).double().mean())
        print(f'{name}.{pn:40s} signflip={sign:.3f} |g|={float(w.norm()):.3e} rel={rel:.2e} floor={floor:.2e} ratio={rel/max(floor,1e-12):.1f}')
