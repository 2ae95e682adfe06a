"""One tiny render per pipeline kind (for compute-sanitizer runs): nerf, append, append_full, smpl; coarse-only and 32+64 too."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_oracle as O          # weights / inputs only (developer tool)
from smpl_nerf_b200 import engine, scene
from tests import helpers as H
for kind, nc, nf, run_fine in (('nerf', 64, 128, 1), ('append', 32, 64, 1), ('append_full', 64, 128, 1), ('smpl', 64, 128, 0), ('smpl', 32, 32, 1)):
    nets = O.build_nets(kind, 1, 'dense')
    args = O.make_args(number_fine_samples=nf, run_fine=run_fine)
    data = scene.data_list(scene.make_rays(5, 5, nc, seed=2), kind)
    gnets, gdata = H.to_cuda(nets, data)
    c, f, w, pe, de, he = gnets
    out = engine.render(kind, c, f, w, args, pe, de, he, gdata)
    torch.cuda.synchronize()
    assert torch.isfinite(out['rgb_fine']).all()
    print(kind, nc, nf, run_fine, 'ok')
