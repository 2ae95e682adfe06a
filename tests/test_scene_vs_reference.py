"""Pins the synthetic-scene ray builder (smpl_nerf_b200/scene.py: the host restatement of utils.get_rays and
datasets/transforms.CoarseSampling that also defines what nrf_generate_rays must produce) bit-for-bit against the
imported reference.  Build container only (skips where /root/reference is absent)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import ref_import as R
from smpl_nerf_b200 import scene

pytestmark = pytest.mark.skipif(not R.available(), reason='reference tree not present')


@pytest.mark.parametrize('hw', [(16, 16), (12, 20)])
def test_camera_rays_match_get_rays(hw):
    h, w = hw
    ref = R.load()
    pose = scene.sphere_pose(12., 40.)
    focal = .5 * w / np.tan(.5 * scene.CAMERA_ANGLE_X)
    t, d = ref.utils.get_rays(h, w, focal, pose)          # utils.py:26-54
    o2, d2 = scene.camera_rays(h, w, pose)
    assert np.array_equal(d.reshape(-1, 3), d2)
    assert np.array_equal(np.broadcast_to(t, d.shape).reshape(-1, 3), o2)


def test_coarse_sampling_matches_transforms():
    spec = importlib.util.spec_from_file_location('ref_transforms', os.path.join(R.REFERENCE_ROOT, 'datasets', 'transforms.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n_rays, nc = 37, 64
    pose = scene.sphere_pose(5., 111.)
    origin, direction = scene.camera_rays(6, 7, pose)
    origin, direction = origin[:n_rays], direction[:n_rays]
    cs, tt = mod.CoarseSampling(scene.NEAR, scene.FAR, nc), mod.ToTensor()
    np.random.seed(7)
    want = [tt(cs((origin[r], direction[r], np.zeros(3, np.float32)))) for r in range(n_rays)]      # one rand() per ray
    z = scene.coarse_depths(n_rays, nc, np.random.RandomState(7))
    pts = origin[:, None, :] + direction[:, None, :] * z[:, :, None]
    for r in range(n_rays):
        assert torch.equal(want[r][0], torch.from_numpy(pts[r]).float())
        assert torch.equal(want[r][3], torch.from_numpy(z[r]).float())
