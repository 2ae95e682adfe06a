import sys, torch
sys.path.insert(0, '/root/repo')
from tests import helpers as H
from tests import test_gpu_train as T
from oracle import nerf_oracle as O
from smpl_nerf_b200 import engine
for scale in (1e3, 1e5, 1e6, 1e8):
    nets = O.build_nets('nerf', 13, 'dense')
    with torch.no_grad():
        for net in nets[:2]:
            net.positions_pose_input.weight.mul_(scale); net.positions_pose_input.bias.mul_(scale)
            net.positional_net[0].weight.div_(scale)
    args = O.make_args()
    data = T._rays('nerf', 6, 6, 64, 2)
    with torch.no_grad():
        want = H.run_oracle('nerf', nets, args, data)
        x = nets[3].encode(data[0])
        amax = float(torch.relu(nets[0].positions_pose_input(x)).max())
    gnets, gdata = H.to_cuda(nets, data)
    with torch.no_grad():
        par = engine.render('nerf', gnets[0], gnets[1], None, args, gnets[3], gnets[4], None, gdata, z_all_in=want['z_all'].to('cuda:0'), taps=True)
        torch.cuda.synchronize()
    s = par['raw_coarse'][..., 3].cpu()
    print(f'scale {scale:g}: max act {amax:.3g} status {int(par["status"].item())} sigma err {float((s - want["raw_coarse"][..., 3]).abs().max()):.3g} '
          f'nan {int(torch.isnan(s).sum())} max|sigma| {float(want["raw_coarse"][..., 3].abs().max()):.3g} w0max {float(nets[0].positions_pose_input.weight.abs().max()):.3g}')
