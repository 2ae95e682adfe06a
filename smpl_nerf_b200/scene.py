"""Synthetic SMPL-NeRF ray batches (host side, numpy float64 -> fp32 like the reference datasets).

The licensed SMPL model, pyrender and smplx are not available offline, so benchmarks and tests
use synthetic rays whose *distribution* mirrors what the reference's dataset code feeds the
pipelines:

  * camera on a sphere of radius 2.4 looking at the origin      (camera.py:86-110, create_dataset.py:30)
  * pin-hole rays, focal = .5 W / tan(.5 * pi/3)                 (utils.py:26-54, smpl_nerf_dataset.py:58)
  * coarse depths linear in disparity between near=1 and far=4, stratified with ONE uniform jitter
    scalar per ray, points computed in float64 then rounded once to fp32   (datasets/transforms.py:82-89,15)
  * goal_pose[69] zero except columns 38 and 41 = arm angle in radians      (render.py:190-220)

Ground-truth colours come from a small analytic "capsule humanoid" so PSNR-vs-GT is defined
without a mesh renderer.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

CAMERA_ANGLE_X = math.pi / 3.0      # create_dataset.py:141
CAMERA_RADIUS = 2.4                 # create_dataset.py:30
NEAR, FAR = 1.0, 4.0                # config_parser.py:68-69


def _rot_xyz_deg(phi: float, theta: float, psi: float) -> np.ndarray:
    """Extrinsic x-then-y-then-z Euler rotation in degrees (what scipy calls 'xyz')."""
    a, b, c = np.radians([phi, theta, psi])
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return rz @ ry @ rx


def sphere_pose(phi: float, theta: float, r: float = CAMERA_RADIUS) -> np.ndarray:
    """4x4 camera-to-world matrix on the sphere (phi: elevation, theta: azimuth, degrees)."""
    pose = np.eye(4)
    pose[:3, :3] = _rot_xyz_deg(-phi, theta, 0.0)
    pose[:3, 3] = [r * np.cos(np.radians(phi)) * np.sin(np.radians(theta)),
                   r * np.sin(np.radians(phi)),
                   r * np.cos(np.radians(phi)) * np.cos(np.radians(theta))]
    return pose


def camera_rays(h: int, w: int, pose: np.ndarray, camera_angle_x: float = CAMERA_ANGLE_X):
    """Per-pixel ray origins [h*w,3] and (un-normalised) directions [h*w,3], row-major pixels."""
    focal = .5 * w / np.tan(.5 * camera_angle_x)
    i, j = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing='xy')
    local = np.stack([(i - w * .5) / focal, -(j - h * .5) / focal, -np.ones_like(i)], -1)
    direction = np.sum(local.reshape(-1, 3)[..., np.newaxis, :] * pose[:3, :3], -1)     # utils.py:52 (elementwise products, then the sum)
    origin = np.broadcast_to(pose[:3, 3], direction.shape)
    return origin, direction


def coarse_bins(n_samples: int, near: float = NEAR, far: float = FAR):
    """(lower, upper) edges of the stratified bins, linear in disparity (datasets/transforms.py:82-86)."""
    t = np.linspace(0., 1., n_samples)
    z = 1. / (1. / near * (1. - t) + 1. / far * t)
    mids = .5 * (z[1:] + z[:-1])
    upper = np.concatenate([mids, z[-1:]])
    lower = np.concatenate([z[:1], mids])
    return lower, upper


def coarse_depths(n_rays: int, n_samples: int, rng: np.random.RandomState,
                  near: float = NEAR, far: float = FAR, jitter: Optional[np.ndarray] = None) -> np.ndarray:
    """[n_rays, n_samples] stratified depths, linear in disparity, one jitter scalar per ray."""
    lower, upper = coarse_bins(n_samples, near, far)
    if jitter is None:
        jitter = rng.rand(n_rays)
    return lower[None, :] + (upper - lower)[None, :] * jitter[:, None]


def _capsule_hit(o, d, a, b, radius):
    """Ray/capsule test (vectorised, returns hit mask and a cheap shading term)."""
    ba = b - a
    oa = o - a
    baba = ba @ ba
    bard = d @ ba
    baoa = oa @ ba
    rdoa = np.einsum('ij,ij->i', d, oa)
    oaoa = np.einsum('ij,ij->i', oa, oa)
    dd = np.einsum('ij,ij->i', d, d)
    A = baba * dd - bard * bard
    B = baba * rdoa - baoa * bard
    C = baba * oaoa - baoa * baoa - radius * radius * baba
    disc = B * B - A * C
    ok = disc >= 0
    t = np.where(ok, (-B - np.sqrt(np.maximum(disc, 0))) / np.maximum(A, 1e-12), np.inf)
    y = baoa + t * bard
    body = ok & (y > 0) & (y < baba) & (t > 0)
    # end caps (spheres)
    hit = body.copy()
    for c in (a, b):
        oc = o - c
        bq = np.einsum('ij,ij->i', d, oc)
        cq = np.einsum('ij,ij->i', oc, oc) - radius * radius
        h = bq * bq - dd * cq
        tt = np.where(h >= 0, (-bq - np.sqrt(np.maximum(h, 0))) / dd, np.inf)
        hit |= (h >= 0) & (tt > 0)
        t = np.minimum(t, np.where((h >= 0) & (tt > 0), tt, np.inf))
    return hit, t


def analytic_colours(origin: np.ndarray, direction: np.ndarray, arm_angle_rad: float) -> np.ndarray:
    """White background, a torso/head/legs/arms capsule figure whose arms lift with the pose."""
    n = origin.shape[0]
    rgb = np.ones((n, 3))
    depth = np.full(n, np.inf)
    s, c = math.sin(arm_angle_rad), math.cos(arm_angle_rad)
    parts = [  # (a, b, radius, colour)
        ((0, -0.1, 0), (0, 0.45, 0), 0.17, (0.8, 0.3, 0.25)),            # torso
        ((0, 0.68, 0), (0, 0.72, 0), 0.12, (0.9, 0.75, 0.6)),             # head
        ((-0.1, -0.2, 0), (-0.12, -0.95, 0), 0.08, (0.2, 0.3, 0.7)),      # legs
        ((0.1, -0.2, 0), (0.12, -0.95, 0), 0.08, (0.2, 0.3, 0.7)),
        ((-0.2, 0.42, 0), (-0.2 - 0.55 * c, 0.42 + 0.55 * s, 0), 0.06, (0.9, 0.75, 0.6)),   # arms
        ((0.2, 0.42, 0), (0.2 + 0.55 * c, 0.42 + 0.55 * s, 0), 0.06, (0.9, 0.75, 0.6)),
    ]
    for a, b, r, col in parts:
        hit, t = _capsule_hit(origin, direction, np.asarray(a, float), np.asarray(b, float), r)
        closer = hit & (t < depth)
        shade = np.clip(1.15 - 0.25 * (t - 1.6), 0.4, 1.0)
        rgb[closer] = np.asarray(col)[None, :] * shade[closer, None]
        depth[closer] = t[closer]
    return rgb


def make_rays(h: int, w: int, n_coarse: int = 64, *, phi: float = 10.0, theta: float = 30.0,
              arm_angle_deg: float = 30.0, seed: int = 0, with_colours: bool = False,
              rng: Optional[np.random.RandomState] = None) -> Dict[str, torch.Tensor]:
    """One view of the synthetic scene as the per-ray tensors a pipeline ``data`` list holds.

    Returns CPU fp32 tensors: ray_samples [B,Nc,3], ray_translation [B,3], ray_direction [B,3],
    z_vals [B,Nc], goal_pose [B,69], rgb [B,3]  (B = h*w, row-major pixels).
    """
    rng = rng or np.random.RandomState(seed)
    pose = sphere_pose(phi, theta)
    origin, direction = camera_rays(h, w, pose)
    z = coarse_depths(origin.shape[0], n_coarse, rng)
    samples = origin[:, None, :] + direction[:, None, :] * z[:, :, None]      # float64, rounded once below
    goal = np.zeros((origin.shape[0], 69))
    goal[:, 38] = goal[:, 41] = np.deg2rad(arm_angle_deg)
    rgb = analytic_colours(origin, direction, np.deg2rad(arm_angle_deg)) if with_colours \
        else np.zeros((origin.shape[0], 3))
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    return dict(ray_samples=f32(samples), ray_translation=f32(origin), ray_direction=f32(direction),
                z_vals=f32(z), goal_pose=f32(goal), rgb=f32(rgb))


def data_list(rays: Dict[str, torch.Tensor], kind: str, sel=slice(None), device=None):
    """Order the tensors like the reference datasets' ``__getitem__`` tuples do."""
    keys = ['ray_samples', 'ray_translation', 'ray_direction', 'z_vals']
    if kind in ('smpl', 'append', 'append_full'):
        keys.append('goal_pose')
    keys.append('rgb')
    out = [rays[k][sel].contiguous() for k in keys]
    if device is not None:
        out = [t.to(device) for t in out]
    return out
