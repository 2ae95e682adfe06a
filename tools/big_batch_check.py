"""One call with 1,048,576 rays (1024 x 1024 view generated on the device): index arithmetic at large offsets.
The tail of the big render must be bit-identical to rendering those rays on their own."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smpl_nerf_b200 import rays, scene
from smpl_nerf_b200.models import SmplNerfPipeline
w = bench.WORKLOADS['cfg2']
coarse, fine, warp, pe, de, he = bench.build_models(w)
dev = torch.device('cuda:0')
pipe = SmplNerfPipeline(coarse.to(dev), fine.to(dev), warp.to(dev), bench.make_args(w), pe, de, he)
side = 1024
data = rays.generate_view(side, side, scene.sphere_pose(10., 30.), n_coarse=64, rng=np.random.RandomState(0))
B = side * side
goal = torch.zeros(B, 69, device=dev); goal[:, 38] = goal[:, 41] = 0.5
data += [goal, torch.zeros(B, 3, device=dev)]
with torch.no_grad():
    out = pipe(data)
    torch.cuda.synchronize()
    assert all(torch.isfinite(o).all() for o in out)
    tail = [t[-1000:].contiguous() for t in data]
    ref = pipe(tail)
    for a, b in zip(out, ref):
        assert torch.equal(a[-1000:], b)
print(f'{B} rays in one call: finite, tail bit-identical to a separate 1000-ray call; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
