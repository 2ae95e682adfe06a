"""Mint the "briefly trained checkpoint" fixture (fixture iv) -- build container only.

    python tests/golden/make_trained.py

Trains a small vanilla NeRF (depth 4, width 256, 32 coarse + 64 fine samples) for a few hundred Adam steps on 32x32
views of the synthetic capsule figure with the REFERENCE classes (imported from /root/reference; loss = MSE(rgb) +
MSE(rgb_fine) as solver/nerf_solver.py:48-51), rounds the weights to fp16-representable values (halves the file; both
sides of every comparison then load the same fp32 numbers), renders a HELD-OUT view with the reference pipeline and
stores: the weights, the held-out rays, the analytic ground truth, the reference's rgb / rgb_fine and its PSNR.
tests/test_gpu_parity.py::test_trained_checkpoint_psnr renders the same view with the engine."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import ref_import as R       # noqa: E402
from smpl_nerf_b200 import scene         # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'trained_nerf_d4.ckpt')
SIDE, NC, NF, STEPS, BATCH = 32, 32, 64, 1200, 512


def main():
    ref = R.load()
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    c, f, _, pe, de, _ = O.build_nets('nerf', 31, 'default', n_layers=4, skips=(), net_cls=ref.RenderRayNet,
                                      enc_cls=ref.PositionalEncoder)
    c.train(); f.train()
    args = O.make_args(number_fine_samples=NF)
    pipe = ref.NerfPipeline(c, f, args, pe, de)
    views = [scene.make_rays(SIDE, SIDE, NC, phi=8.0 + 4 * (k % 3), theta=30.0 * k, arm_angle_deg=30.0, seed=100 + k,
                             with_colours=True) for k in range(12)]
    train = {k: torch.cat([v[k] for v in views]) for k in views[0]}
    opt = torch.optim.Adam(list(c.parameters()) + list(f.parameters()), lr=1e-3)
    n = train['z_vals'].shape[0]
    # 96% of the pixels are white background: draw half of every batch from the figure, or the nets settle on "all white"
    fg = torch.nonzero((train['rgb'] < 0.99).any(-1)).flatten()
    print(f'{n} training rays, {fg.numel()} on the figure')
    t0 = time.time()
    for step in range(STEPS):
        sel = torch.cat([torch.randint(0, n, (BATCH // 2,)), fg[torch.randint(0, fg.numel(), (BATCH // 2,))]])
        data = scene.data_list(train, 'nerf', sel)
        rgb, rgb_fine, _, _ = pipe(data)
        loss = torch.mean((rgb - data[-1]) ** 2) + torch.mean((rgb_fine - data[-1]) ** 2)
        opt.zero_grad(); loss.backward(); opt.step()
        if step % 50 == 0:
            print(f'step {step:4d} loss {float(loss):.5f}  ({time.time() - t0:.0f} s)', flush=True)
    c.eval(); f.eval()
    with torch.no_grad():
        for net in (c, f):
            for p in net.parameters():
                p.copy_(p.half().float())
        held = scene.make_rays(SIDE, SIDE, NC, phi=10.0, theta=45.0, arm_angle_deg=30.0, seed=999, with_colours=True)
        data = scene.data_list(held, 'nerf')
        rgb, rgb_fine, _, alpha = pipe(data)
    mse = float(torch.mean((rgb_fine.double() - data[-1].double()) ** 2))
    psnr = -10.0 * np.log10(mse)
    white = -10.0 * np.log10(float(torch.mean((1.0 - data[-1].double()) ** 2)))
    print(f'held-out PSNR of the reference render vs ground truth: {psnr:.3f} dB (an all-white image scores {white:.3f} dB)')
    assert psnr > white + 3.0, 'the checkpoint did not learn the figure'
    torch.save(dict(coarse={k: v.half() for k, v in c.state_dict().items()}, fine={k: v.half() for k, v in f.state_dict().items()},
                    n_layers=4, skips=(), n_coarse=NC, n_fine=NF, side=SIDE, data=[t.clone() for t in data],
                    reference_rgb=rgb.clone(), reference_rgb_fine=rgb_fine.clone(), reference_alpha=alpha.clone(),
                    reference_psnr=psnr, steps=STEPS, torch_version=torch.__version__), OUT)
    print(f'{OUT}: {os.path.getsize(OUT) / 1024:.0f} KB')


if __name__ == '__main__':
    main()
