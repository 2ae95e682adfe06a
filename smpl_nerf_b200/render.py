"""Full-frame render / evaluation driver: the loop of inference.py:222-265 without
its dataset, PNG and LPIPS dependencies.

    frames = render_frames(pipeline, cameras, poses, h, w)        # [n, h, w, 3] on the device
    print(scores(frames, ground_truth))                           # MSE / PSNR / SSIM (util/scores.py:457-464 minus LPIPS)
    save_frames(frames, 'renders/run')                            # img_000.png ... (inference.py:268-274)

Rays are generated on the device (``rays.generate_view``), each frame is ONE pipeline call (the reference walks a
DataLoader in ``inf_batchsize`` = 800-ray batches, inference.py:231,247-254, because it materialises
[rays x samples x features] tensors; the fused kernel does not), and under ``torch.distributed`` every frame's rays
are sharded over the ranks -- each rank generates only its own rays -- with a single all-gather of the rendered tiles.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

import torch.distributed as tdist

from . import dist as nd
from . import ops, rays


def frame_data(h: int, w: int, camera_transform, goal_pose: Optional[Sequence[float]], n_coarse: int, device=None, *,
               rng: Optional[np.random.RandomState] = None, near: float = rays.scene.NEAR, far: float = rays.scene.FAR,
               camera_angle_x: float = rays.scene.CAMERA_ANGLE_X, ray_range=None) -> List[torch.Tensor]:
    """The reference's per-batch ``data`` list for the rays ``ray_range`` (default: all h*w) of one view, built on the device."""
    data = rays.generate_view(h, w, camera_transform, camera_angle_x=camera_angle_x, near=near, far=far, n_coarse=n_coarse,
                              rng=rng, device=device, ray_range=ray_range)
    B = int(data[0].shape[0])
    dev = data[0].device
    if goal_pose is not None:
        gp = torch.as_tensor(np.asarray(goal_pose, dtype=np.float32), device=dev).reshape(1, -1)
        data.append(gp.expand(B, gp.shape[1]).contiguous())        # datasets/smpl_nerf_dataset.py:63 (pose repeated per ray)
    data.append(torch.zeros(B, 3, dtype=torch.float32, device=dev))      # rgb slot: never read by the pipelines
    return data


def render_frames(pipeline, cameras: Sequence, poses: Optional[Sequence], h: int, w: int, *, n_coarse: int = 64,
                  device=None, seed: Optional[int] = 0, out_index: int = 1) -> torch.Tensor:
    """Render ``len(cameras)`` views; returns ``[n, h, w, 3]`` fp32 images (``out[out_index]`` = rgb_fine, as
    inference.py:252 reads it), identical on every rank when torch.distributed is initialised.  Every rank GENERATES and
    renders only its contiguous shard of each frame's rays; one all-gather per frame assembles the image.
    ``device``: default = the current CUDA device (under torchrun: the rank's own GPU after ``torch.cuda.set_device``)."""
    rng = np.random.RandomState(seed) if seed is not None else None
    world = tdist.get_world_size() if tdist.is_available() and tdist.is_initialized() else 1
    rank = tdist.get_rank() if world > 1 else 0
    frames = []
    with torch.no_grad():
        for k, cam in enumerate(cameras):
            window = nd.shard_range(h * w, rank, world)
            data = frame_data(h, w, cam, None if poses is None else poses[k], n_coarse, device, rng=rng, ray_range=window)
            local = pipeline(data)[out_index]
            frames.append(nd.gather_tiles(local.contiguous(), h * w).reshape(h, w, 3))
    return torch.stack(frames, 0)


def psnr_per_frame(frames: torch.Tensor, ground_truth: torch.Tensor) -> List[float]:
    """-10 log10(mse) per frame: util/scores.py:30-48 img2psnr / utils.py:484-488 mse2psnr."""
    return [nd.psnr(a, b) for a, b in zip(frames, ground_truth.to(frames.device))]


def scores(frames: torch.Tensor, ground_truth: torch.Tensor) -> dict:
    """``print_scores`` of util/scores.py:457-464 without its LPIPS term (which downloads VGG weights): MSE, PSNR and SSIM
    of ``[n, h, w, 3]`` image stacks, all reduced on the device; one host read at the end."""
    x = frames.permute(0, 3, 1, 2).contiguous()                       # inference.py:258
    y = ground_truth.to(frames.device).permute(0, 3, 1, 2).contiguous()
    vals = torch.stack([ops.img2mse(x, y), ops.img2psnr(x, y), ops.ssim(x, y)]).tolist()
    return {'mse': vals[0], 'psnr': vals[1], 'ssim': vals[2]}


def save_frames(frames: torch.Tensor, output_dir: str, prefix: str = 'img_') -> List[str]:
    """inference.py:260-275 ``save_rerenders``: clip / scale / uint8 / BGR on the device, one D2H copy, ``img_%03d.png`` files
    written with OpenCV (the reference uses imageio, absent here; its GIF is not written)."""
    import os

    import cv2
    os.makedirs(output_dir, exist_ok=True)
    bgr = ops.to_uint8_bgr(frames, to_bgr=True).cpu().numpy()
    paths = []
    for i, img in enumerate(bgr):
        path = os.path.join(output_dir, f'{prefix}{i:03d}.png')
        if not cv2.imwrite(path, img):
            raise RuntimeError(f'could not write {path}')
        paths.append(path)
    return paths
