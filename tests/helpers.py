"""Shared helpers for the test-suite (tests may use oracle/ as the checker)."""
import glob
import os

import torch

from oracle import nerf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# Tolerances from BASELINE.json north_star: 1e-3 abs on RGB, 1e-4 abs on sigma (and on alpha, which is
# what the pipelines expose as "densities").
TOL_RGB = 1e-3
TOL_SIGMA = 1e-4
TOL_ALPHA = 1e-4


def fixtures():
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.pt')))


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + '.pt'), weights_only=False)


def nets_for(fx):
    """Rebuild the fixture's nets from its seed and check the weight checksum."""
    nets = O.build_nets(fx['kind'], fx['seed'], fx['variant'], **fx['build'])
    assert O.weight_checksum(list(nets[:3])) == fx['weight_checksum'], \
        'default-init weights are not reproducible from the seed on this torch build'
    return nets


def args_for(fx, **kw):
    return O.make_args(run_fine=fx['run_fine'], human_pose_encoding=1 if fx['pose_encoded'] else 0, **kw)


def run_oracle(kind, nets, args, data, **kw):
    c, f, w, pe, de, he = nets
    if kind == 'nerf':
        return O.nerf_forward(c, f, pe, de, args, data, **kw)
    if kind == 'append':
        return O.append_to_nerf_forward(c, f, pe, de, he, args, data, **kw)
    if kind == 'append_full':
        return O.append_smpl_params_forward(c, f, pe, de, he, args, data, **kw)
    return O.smpl_nerf_forward(c, f, w, pe, de, he, args, data, **kw)


def to_cuda(nets, data, dev='cuda:0'):
    c, f, w, pe, de, he = nets
    import copy
    g = [copy.deepcopy(m).to(dev) if m is not None else None for m in (c, f, w)]
    return (g[0], g[1], g[2], pe, de, he), [t.to(dev) for t in data]


def alpha_mask_well_conditioned(sigma_ref, thresh=1e-3):
    """The last sample's alpha is 1-exp(-relu(sigma)*1e10): a step function of sigma at 0, so any
    implementation (including the reference in another precision) may flip it when |sigma| ~ 0.
    Compare alpha there only where the reference sigma is clearly away from the kink."""
    m = torch.ones_like(sigma_ref, dtype=torch.bool)
    m[..., -1] = sigma_ref[..., -1].abs() > thresh
    return m


def load_trained():
    """The briefly trained vanilla-NeRF checkpoint + held-out view minted by tests/golden/make_trained.py."""
    ck = torch.load(os.path.join(GOLDEN_DIR, 'trained_nerf_d4.ckpt'), weights_only=False)
    c, f, _, pe, de, he = O.build_nets('nerf', 0, 'default', n_layers=ck['n_layers'], skips=tuple(ck['skips']))
    c.load_state_dict({k: v.float() for k, v in ck['coarse'].items()})
    f.load_state_dict({k: v.float() for k, v in ck['fine'].items()})
    args = O.make_args(number_fine_samples=ck['n_fine'])
    return ck, (c, f, None, pe, de, he), args


def load_trained_smpl():
    """The briefly trained SmplNerfPipeline (8 x 256 x 2 + warp net, full fp32 weights) minted by tests/golden/make_trained_smpl.py.
    Returns (ck, nets, args); ``trained_smpl_view(ck, name)`` regenerates a stored evaluation view's rays."""
    ck = torch.load(os.path.join(GOLDEN_DIR, 'trained_smpl_d8.ckpt'), weights_only=False)
    c, f, w, pe, de, he = O.build_nets('smpl', 0, 'default')
    c.load_state_dict(ck['coarse']); f.load_state_dict(ck['fine']); w.load_state_dict(ck['warp'])
    for m in (c, f, w):
        m.eval()
    return ck, (c, f, w, pe, de, he), O.make_args(number_fine_samples=ck['n_fine'])


def trained_smpl_view(ck, name):
    """Rays of evaluation view ``name`` ('seen' / 'heldout'): regenerated from the stored scene.make_rays arguments (numpy float64 +
    RandomState: bit-reproducible) and checked against the stored checksums."""
    from smpl_nerf_b200 import scene
    v = ck['views'][name]
    data = scene.data_list(scene.make_rays(**v['args']), 'smpl')
    assert [float(t.double().sum()) for t in data] == v['data_checksum'], 'the evaluation rays are not reproducible on this machine'
    return v, data


def psnr(img, ref):
    return float(-10.0 * torch.log10(torch.mean((img.double().cpu() - ref.double().cpu()) ** 2)))
