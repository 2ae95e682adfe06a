"""Per-kernel time of one training step (torch.profiler / CUPTI).  Usage: python tools/train_profile.py [batch] [fast]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from smpl_nerf_b200 import scene  # noqa: E402
from smpl_nerf_b200.models import SmplNerfPipeline  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = bench.WORKLOADS['train']
dev = torch.device('cuda:0')
coarse, fine, warp, pe, de, he = bench.build_models(w)
nets = [m.to(dev).train() for m in (coarse, fine, warp)]
pipe = SmplNerfPipeline(nets[0], nets[1], nets[2], bench.make_args(w), pe, de, he)
pipe.precision = 1 if 'fast' in sys.argv else 0
params = [p for m in nets for p in m.parameters()]
opt = torch.optim.Adam(params, lr=5e-4, fused=True)
rays = scene.make_rays(64, 64, 64, seed=7, with_colours=True)
data = [t[:B].to(dev) for t in scene.data_list(rays, 'smpl')]


def step():
    out = pipe(data)
    loss = torch.mean((out[0] - data[-1]) ** 2) + torch.mean((out[1] - data[-1]) ** 2)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
