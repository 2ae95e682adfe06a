"""Full-frame render / evaluation driver: the loop of inference.py:222-265 without
its dataset, PNG and LPIPS dependencies.

    frames = render_frames(pipeline, cameras, poses, h, w)        # [n, h, w, 3] on the device
    scores = psnr_per_frame(frames, ground_truth)

Rays are generated on the device (``rays.generate_view``), each frame is ONE pipeline call (the reference walks a
DataLoader in ``inf_batchsize`` = 800-ray batches, inference.py:231,247-254, because it materialises
[rays x samples x features] tensors; the fused kernel does not), and under ``torch.distributed`` every frame's rays
are sharded over the ranks with a single all-gather of the rendered tiles (``dist.render_frame_sharded``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import dist as nd
from . import rays


def frame_data(h: int, w: int, camera_transform, goal_pose: Optional[Sequence[float]], n_coarse: int, device, *,
               rng: Optional[np.random.RandomState] = None, near: float = rays.scene.NEAR, far: float = rays.scene.FAR,
               camera_angle_x: float = rays.scene.CAMERA_ANGLE_X) -> List[torch.Tensor]:
    """The reference's per-batch ``data`` list for all h*w rays of one view, built on the device."""
    data = rays.generate_view(h, w, camera_transform, camera_angle_x=camera_angle_x, near=near, far=far, n_coarse=n_coarse,
                              rng=rng, device=device)
    B = h * w
    if goal_pose is not None:
        gp = torch.as_tensor(np.asarray(goal_pose, dtype=np.float32), device=device).reshape(1, -1)
        data.append(gp.expand(B, gp.shape[1]).contiguous())        # datasets/smpl_nerf_dataset.py:63 (pose repeated per ray)
    data.append(torch.zeros(B, 3, dtype=torch.float32, device=device))      # rgb slot: never read by the pipelines
    return data


def render_frames(pipeline, cameras: Sequence, poses: Optional[Sequence], h: int, w: int, *, n_coarse: int = 64,
                  device='cuda:0', seed: Optional[int] = 0, out_index: int = 1) -> torch.Tensor:
    """Render ``len(cameras)`` views; returns ``[n, h, w, 3]`` fp32 images (``out[out_index]`` = rgb_fine, as
    inference.py:252 reads it), identical on every rank when torch.distributed is initialised."""
    rng = np.random.RandomState(seed) if seed is not None else None
    frames = []
    with torch.no_grad():
        for k, cam in enumerate(cameras):
            data = frame_data(h, w, cam, None if poses is None else poses[k], n_coarse, device, rng=rng)
            img = nd.render_frame_sharded(pipeline, data, out_index=out_index)
            frames.append(img.reshape(h, w, 3))
    return torch.stack(frames, 0)


def psnr_per_frame(frames: torch.Tensor, ground_truth: torch.Tensor) -> List[float]:
    """-10 log10(mse) per frame: util/scores.py:30-48 img2psnr / utils.py:484-488 mse2psnr."""
    return [nd.psnr(a, b) for a, b in zip(frames, ground_truth.to(frames.device))]
