"""Pins the oracle against the committed golden fixtures (minted from the reference itself by
tests/golden/make_golden.py).  Runs everywhere, including the GPU box where /root/reference is absent."""
import pytest
import torch

from oracle import nerf_oracle as O
from tests import helpers as H


def test_fixtures_present():
    names = H.fixtures()
    assert len(names) >= 10
    kinds = {H.load_fixture(n)['kind'] for n in names}
    assert kinds == {'nerf', 'append', 'append_full', 'smpl'}


@pytest.mark.parametrize('name', H.fixtures())
def test_oracle_reproduces_reference_outputs(name):
    fx = H.load_fixture(name)
    nets = H.nets_for(fx)
    args = H.args_for(fx)
    with torch.no_grad():
        out = H.run_oracle(fx['kind'], nets, args, fx['data'])
    got = O.as_tuple(out)
    want = fx['reference_outputs']
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert torch.equal(a, b), f'{name}: oracle deviates from the reference fixture by {(a - b).abs().max()}'
    for k, v in fx['intermediates'].items():
        assert torch.equal(out[k], v), k


@pytest.mark.parametrize('name', ['nerf_dense', 'smpl_dense'])
def test_fp64_noise_floor_is_small_but_nonzero(name):
    """Documents the reference's own fp32-vs-fp64 deviation (context for the parity tolerances)."""
    fx = H.load_fixture(name)
    d = [float((a.double() - b).abs().max()) for a, b in zip(fx['reference_outputs'], fx['reference_outputs_fp64'])]
    assert max(d[:2]) < 1e-5            # colours
    assert 0 < max(d) < 1e-2            # per-sample outputs move more (ill-conditioned sampler)


def test_analytic_kats():
    """Known-answer tests of the compositing / sampling restatement."""
    B, n = 3, 64
    z = torch.linspace(1, 4, n).expand(B, n).contiguous()
    dirs = torch.zeros(B, n, 3); dirs[..., 2] = 1
    raw = torch.randn(B, n, 4)
    raw[..., 3] = -raw[..., 3].abs()                      # sigma <= 0 everywhere -> alpha = 0
    rgb, w, a = O.composite(raw, z, dirs, white_background=0)
    assert a.abs().max() == 0 and rgb.abs().max() == 0
    rgb, w, a = O.composite(raw, z, dirs, white_background=1)
    assert torch.equal(rgb, torch.ones(B, 3))
    raw[..., 3] = -1.
    raw[:, 17, 3] = 1e6                                    # one opaque sample -> one-hot weights
    rgb, w, a = O.composite(raw, z, dirs, white_background=0)
    assert torch.allclose(w[:, 17], torch.ones(B)) and w.sum(-1).allclose(torch.ones(B))
    assert torch.allclose(rgb, torch.sigmoid(raw[:, 17, :3]))
    bins = torch.linspace(1, 4, n - 1).expand(B, n - 1).contiguous()     # uniform weights -> linear samples
    zs = O.inverse_cdf(bins, torch.ones(B, n - 2), 128)
    u = torch.linspace(0, 1, 128)
    assert torch.allclose(zs, (bins[:, :1] + u * (bins[:, -1:] - bins[:, :1])), atol=2e-6)


def test_oracle_reproduces_the_trained_checkpoint_render():
    """Held-out view of the briefly trained checkpoint: the oracle equals the reference render bit for bit."""
    ck, nets, args = H.load_trained()
    with torch.no_grad():
        out = H.run_oracle('nerf', nets, args, ck['data'])
    assert torch.equal(out['rgb'], ck['reference_rgb']) and torch.equal(out['rgb_fine'], ck['reference_rgb_fine'])
    assert torch.equal(out['alpha_out'], ck['reference_alpha'])
    assert abs(H.psnr(out['rgb_fine'], ck['data'][-1]) - ck['reference_psnr']) < 1e-9
