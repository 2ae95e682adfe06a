"""On-device ray generation + coarse sampling for one camera view.

Replaces, for full-frame rendering, the host pipeline ``utils.get_rays`` (utils.py:26-54) ->
``SmplNerfDataset.__getitem__`` (datasets/smpl_nerf_dataset.py:95-101) -> ``CoarseSampling`` + ``ToTensor``
(datasets/transforms.py:82-89, 13-19) -> collate -> ``.to(device)``: instead of ~1.3 KB per ray crossing PCIe, the
host ships a 4x4 camera matrix, two n_coarse-long bin tables and ONE jitter scalar per ray, and
``nrf_generate_rays`` builds ``ray_samples / ray_translation / ray_direction / z_vals`` in HBM with the reference's
float64-then-cast arithmetic (bit-identical to ``scene.make_rays`` -- tests/test_gpu_ops.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib, scene
from ._lib import check

_bins_cache = {}


def _bins(n_coarse: int, near: float, far: float, device):
    key = (n_coarse, float(near), float(far), str(device))
    if key not in _bins_cache:
        lower, upper = scene.coarse_bins(n_coarse, near, far)
        _bins_cache[key] = (torch.from_numpy(lower).to(device), torch.from_numpy(upper - lower).to(device))
    return _bins_cache[key]


def generate_view(h: int, w: int, camera_transform: np.ndarray, *, camera_angle_x: float = scene.CAMERA_ANGLE_X,
                  near: float = scene.NEAR, far: float = scene.FAR, n_coarse: int = 64, jitter=None,
                  rng: Optional[np.random.RandomState] = None, device=None, ray_range=None) -> List[torch.Tensor]:
    """-> ``[ray_samples[B,Nc,3], ray_translation[B,3], ray_direction[B,3], z_vals[B,Nc]]`` (fp32, on ``device``),
    B = h*w rays in row-major pixel order.  ``jitter``: [B] float64 (host array or device tensor) -- the one
    ``np.random.rand()`` scalar CoarseSampling draws per ray; drawn from ``rng`` (default: numpy's global stream,
    like the reference) when omitted.  ``ray_range = (start, stop)``: build only those rays of the view (a rank's shard of a
    multi-GPU render); the jitter stream is still drawn for the WHOLE view so every rank sees the same scalars.
    ``device``: default = the current CUDA device."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('smpl_nerf_b200.rays runs on CUDA devices only (no CPU fallback)')
    cam = np.ascontiguousarray(np.asarray(camera_transform, dtype=np.float64))
    if cam.shape != (4, 4):
        raise ValueError(f'camera_transform must be 4x4, got {cam.shape}')
    B = h * w
    if jitter is None:
        jitter = (rng.rand(B) if rng is not None else np.random.rand(B))
    if not isinstance(jitter, torch.Tensor):
        jitter = torch.from_numpy(np.ascontiguousarray(np.asarray(jitter, dtype=np.float64)))
    if jitter.numel() != B:
        raise ValueError(f'jitter must have {B} entries, got {jitter.numel()}')
    r0, r1 = (0, B) if ray_range is None else (int(ray_range[0]), int(ray_range[1]))
    if not (0 <= r0 <= r1 <= B):
        raise ValueError(f'ray_range {ray_range} outside [0, {B}]')
    jitter = jitter.reshape(-1)[r0:r1].to(device=device, dtype=torch.float64).contiguous()      # only the shard crosses PCIe
    B = r1 - r0
    focal = float(.5 * w / np.tan(.5 * camera_angle_x))          # datasets/smpl_nerf_dataset.py:58
    with torch.cuda.device(device):
        lower, span = _bins(n_coarse, near, far, device)
        samples = torch.empty(B, n_coarse, 3, dtype=torch.float32, device=device)
        origin = torch.empty(B, 3, dtype=torch.float32, device=device)
        direction = torch.empty(B, 3, dtype=torch.float32, device=device)
        z = torch.empty(B, n_coarse, dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(_lib.lib().nrf_generate_rays_range(h, w, focal, cam.ctypes.data_as(C.POINTER(C.c_double)), lower.data_ptr(), span.data_ptr(),
                                                 jitter.data_ptr(), n_coarse, r0, B, samples.data_ptr(), origin.data_ptr(),
                                                 direction.data_ptr(), z.data_ptr(), stream), 'nrf_generate_rays_range')
        jitter.record_stream(torch.cuda.current_stream(device))
    return [samples, origin, direction, z]
