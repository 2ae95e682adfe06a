"""Stand-alone ops of the C ABI against the oracle / numpy (GPU)."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from smpl_nerf_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('L,ident,c', [(10, False, 3), (4, False, 3), (10, False, 2), (3, True, 3), (0, True, 3)])
def test_positional_encoding(L, ident, c):
    torch.manual_seed(0)
    x = torch.randn(37, 5, c) * 2
    want = O.Encoder(L, ident).encode(x)
    got = ops.PositionalEncoder(L, ident).encode(x.to(DEV)).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6


@pytest.mark.parametrize('white', [0, 1])
def test_raw2outputs(white):
    torch.manual_seed(1)
    B, n = 50, 192
    raw = torch.randn(B, n, 4)
    z = torch.sort(torch.rand(B, n) * 3 + 1, -1)[0]
    dirs = torch.randn(B, n, 3)
    want = O.composite(raw, z, dirs, white_background=white)
    args = O.make_args(white_background=white)
    got = ops.raw2outputs(raw.to(DEV), z.to(DEV), dirs.to(DEV), args)
    for a, b, tol in zip(got, want, (1e-5, 1e-6, 1e-6)):
        assert float((a.cpu() - b).abs().max()) <= tol


def test_sample_pdf():
    torch.manual_seed(2)
    B = 64
    z = torch.sort(torch.rand(B, 64) * 3 + 1, -1)[0]
    bins = .5 * (z[:, 1:] + z[:, :-1])
    w = torch.rand(B, 62) ** 6
    w[0] = 0
    w[1, 5:] = 0
    want = O.inverse_cdf(bins, w, 128)
    got = ops.sample_pdf(bins.to(DEV), w.to(DEV), O.make_args()).cpu()
    err = (got - want).abs()
    # The reference's sampler is a step function of the CDF in two places: utils.py:224 switches denom
    # to 1 when cdf[above]-cdf[below] < 1e-5, and searchsorted flips its index when u hits a CDF entry
    # (u = 1 vs cdf[-1] ~ 1 is the usual case) -- an ulp of difference in the pdf normalisation moves
    # the result by up to a bin there (the reference's own fp32 vs fp64 runs disagree).  Compare
    # strictly where the fp64 CDF is clear of both switches, and bound the rest by the bin width.
    wd = w.double() + 1e-5
    cdf = torch.cat([torch.zeros(B, 1, dtype=torch.float64), torch.cumsum(wd / wd.sum(-1, keepdim=True), -1)], -1)
    u = torch.linspace(0., 1., 128, dtype=torch.float32).double().expand(B, 128).contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    denom = torch.gather(cdf, -1, idx.clamp(max=62)) - torch.gather(cdf, -1, (idx - 1).clamp(min=0))
    near_entry = (u[:, :, None] - cdf[:, None, :]).abs().min(-1)[0] < 1e-6
    clear = ((denom - 1e-5).abs() > 1e-6) & ~near_entry
    assert float(clear.float().mean()) > 0.95
    assert float(err[clear].max()) <= 1e-3     # (u - c0) / denom amplifies an ulp of c0 by up to 1/1.1e-5
    assert float(torch.quantile(err, 0.99)) <= 1e-5
    assert float(err.max()) <= float((bins[:, 1:] - bins[:, :-1]).max()) + 1e-5


@pytest.mark.parametrize('nc,nf', [(64, 128), (32, 64), (48, 80), (17, 5)])
def test_fine_sampling(nc, nf):
    """ops.fine_sampling (nrf_fine_sampling, utils.py:231-264) against the oracle: the merged depths are sorted, hold the
    coarse depths bit-for-bit, the points are o + d * z with the reference's separate multiply and add, and the new
    samples agree with the oracle's away from the sampler's two step-function switches (see test_sample_pdf)."""
    torch.manual_seed(5)
    B = 96
    z = torch.sort(torch.rand(B, nc) * 3 + 1, -1)[0]
    w = torch.rand(B, nc) ** 4
    w[0] = 0
    o, d = torch.randn(B, 3), torch.randn(B, 3)
    args = O.make_args(number_fine_samples=nf)
    z_want, pts_want, z_new = O.fine_samples(o, d, z, w, nf)
    z_got, pts_got = ops.fine_sampling(o.to(DEV), d.to(DEV), z.to(DEV), w.to(DEV), args)
    z_got, pts_got = z_got.cpu(), pts_got.cpu()
    assert z_got.shape == (B, nc + nf) and pts_got.shape == (B, nc + nf, 3)
    assert bool((z_got[:, 1:] >= z_got[:, :-1]).all())
    # points are exactly o + d * z of the kernel's own depths (separate fp32 multiply and add, utils.py:262)
    assert torch.equal(pts_got, o[:, None, :] + d[:, None, :] * z_got[:, :, None])
    # every coarse depth is present bit-for-bit
    for b in range(0, B, 7):
        assert set(z[b].tolist()) <= set(z_got[b].tolist())
    err = (z_got - z_want).abs()
    assert float(torch.quantile(err.flatten(), 0.98)) <= 1e-5
    assert float(err.max()) <= float((z[:, 1:] - z[:, :-1]).max()) + 1e-5
    # where the kernel's depths agree with the oracle's to a few ulp the points do too
    close = err <= 2e-6
    assert float(close.float().mean()) > 0.9
    assert float((pts_got[close] - pts_want[close]).abs().max()) <= 2e-5


@pytest.mark.parametrize('Ba,Bv', [(1, 1), (100, 100), (1, 100), (100, 1)])
@pytest.mark.parametrize('A,V', [(1, 1), (50, 12), (500, 120), (63, 128)])
@pytest.mark.parametrize('side', ['left', 'right'])
def test_searchsorted_matches_numpy(Ba, Bv, A, V, side):
    """Same grid as torchsearchsorted/test/test_searchsorted.py:34-44 (numpy row-wise oracle), seeded."""
    rng = np.random.RandomState(A * 1000 + V + Ba)
    a = np.sort(rng.rand(Ba, A).astype(np.float32), -1)
    v = rng.rand(Bv, V).astype(np.float32)
    got = ops.searchsorted(torch.from_numpy(a).to(DEV), torch.from_numpy(v).to(DEV), side=side)
    assert got.dtype == torch.int64 and tuple(got.shape) == (max(Ba, Bv), V)
    want = np.stack([np.searchsorted(a[i if Ba > 1 else 0], v[i if Bv > 1 else 0], side=side) for i in range(max(Ba, Bv))])
    assert np.array_equal(got.cpu().numpy(), want)


def test_searchsorted_ties_and_out_argument():
    a = torch.tensor([[0., 1., 1., 1., 2., 3.]], device=DEV)
    v = torch.tensor([[-1., 0., 1., 1.5, 3., 4.]], device=DEV)
    assert ops.searchsorted(a, v, side='left').tolist() == [[0, 0, 1, 4, 5, 6]]
    assert ops.searchsorted(a, v, side='right').tolist() == [[0, 1, 4, 4, 6, 6]]
    out = torch.empty(1, 6, dtype=torch.long, device=DEV)
    res = ops.searchsorted(a, v, out=out, side='right')
    assert res is out
    with pytest.raises(AssertionError):
        ops.searchsorted(a, v, out=torch.empty(1, 6, dtype=torch.int32, device=DEV))
    with pytest.raises(AssertionError):
        ops.searchsorted(a[0], v)


@pytest.mark.parametrize('hw,nc', [((32, 32), 64), ((17, 23), 32)])
def test_generate_rays_is_bit_identical_to_the_host_builder(hw, nc):
    """nrf_generate_rays (float64 on the device) against scene.make_rays (numpy float64, pinned to the reference's
    get_rays / CoarseSampling by tests/test_scene_vs_reference.py): every output bit-equal."""
    import numpy as np
    from smpl_nerf_b200 import rays, scene
    h, w = hw
    want = scene.make_rays(h, w, nc, phi=7.0, theta=123.0, seed=11)
    jitter = np.random.RandomState(11).rand(h * w)
    got = rays.generate_view(h, w, scene.sphere_pose(7.0, 123.0), n_coarse=nc, jitter=jitter)
    torch.cuda.synchronize()
    for g, k in zip(got, ['ray_samples', 'ray_translation', 'ray_direction', 'z_vals']):
        assert g.dtype == torch.float32 and g.shape == want[k].shape
        assert torch.equal(g.cpu(), want[k]), k


def test_render_driver_matches_batched_pipeline_calls():
    """render.render_frames (device ray generation, one call per frame) == the pipeline fed host-built rays."""
    import numpy as np
    from oracle import nerf_oracle as O
    from smpl_nerf_b200 import render, scene
    from smpl_nerf_b200.models import SmplNerfPipeline
    from tests import helpers as H
    nets = O.build_nets('smpl', 3, 'dense')
    args = O.make_args()
    (c, f, wn, pe, de, he), _ = H.to_cuda(nets, [])
    pipe = SmplNerfPipeline(c, f, wn, args, pe, de, he)
    h = w = 12
    cams = [scene.sphere_pose(10.0, 30.0), scene.sphere_pose(-5.0, 200.0)]
    goal = np.zeros(69, np.float32); goal[38] = goal[41] = np.deg2rad(25.0)
    frames = render.render_frames(pipe, cams, [goal, goal], h, w, seed=5)
    assert frames.shape == (2, h, w, 3)
    rng = np.random.RandomState(5)
    for k, (phi, theta) in enumerate([(10.0, 30.0), (-5.0, 200.0)]):
        rays_h = scene.make_rays(h, w, 64, phi=phi, theta=theta, arm_angle_deg=25.0, rng=rng)
        data = [t.cuda() for t in scene.data_list(rays_h, 'smpl')]
        with torch.no_grad():
            want = pipe(data)[1]
        assert torch.equal(frames[k].reshape(-1, 3), want)
    psnr = render.psnr_per_frame(frames, frames.clone() + 0.1)
    assert all(abs(p - 20.0) < 1e-3 for p in psnr)
    # a rank's window of the view == the same rows of the whole view (what a sharded render generates)
    import numpy as np
    from smpl_nerf_b200 import rays
    jit = np.random.RandomState(3).rand(h * w)
    full = rays.generate_view(h, w, cams[0], n_coarse=64, jitter=jit)
    part = rays.generate_view(h, w, cams[0], n_coarse=64, jitter=jit, ray_range=(37, 101))
    for a, b in zip(full, part):
        assert torch.equal(a[37:101], b)


@pytest.mark.parametrize('shape,ks', [((2, 3, 64, 64), 11), ((1, 3, 37, 53), 11), ((3, 1, 16, 128), 7), ((1, 3, 11, 11), 11)])
def test_ssim_matches_the_reference_formula(shape, ks):
    """ops.ssim (nrf_ssim) against the oracle restatement of util/scores.py:88-173 (pinned to the reference on the CPU)."""
    from oracle import scores_oracle as S
    torch.manual_seed(sum(shape))
    x = torch.rand(shape)
    y = (x + 0.1 * torch.randn(shape)).clamp(0, 1)
    want, want_cs = S.ssim(x, y, kernel_size=ks, full=True)
    got, got_cs = ops.ssim(x.to(DEV), y.to(DEV), kernel_size=ks, full=True)
    assert abs(float(got) - float(want)) <= 2e-6 and abs(float(got_cs) - float(want_cs)) <= 2e-6
    none = ops.ssim(x.to(DEV), y.to(DEV), kernel_size=ks, reduction='none')
    assert none.shape == (shape[0],) and float((none.cpu() - S.ssim(x, y, kernel_size=ks, reduction='none')).abs().max()) <= 2e-6
    assert abs(float(ops.ssim(x.to(DEV), x.to(DEV), kernel_size=ks)) - 1.0) <= 1e-6
    assert abs(float(ops.img2psnr(x.to(DEV), y.to(DEV))) - float(S.psnr(x, y))) <= 1e-4
    with pytest.raises(ValueError):
        ops.ssim(x.to(DEV)[..., :5, :5], y.to(DEV)[..., :5, :5], kernel_size=ks)


def test_scores_and_png_writer(tmp_path):
    """render.scores (MSE / PSNR / SSIM on the device) and render.save_frames (inference.py:260-274: clip, x255, uint8, BGR, PNG)."""
    import cv2
    import numpy as np
    from oracle import scores_oracle as S
    from smpl_nerf_b200 import render
    torch.manual_seed(3)
    frames = torch.rand(2, 24, 20, 3) * 1.2 - 0.1
    truth = frames.clamp(0, 1) * 0.9
    sc = render.scores(frames.to(DEV), truth.to(DEV))
    x, y = frames.permute(0, 3, 1, 2), truth.permute(0, 3, 1, 2)
    assert abs(sc['mse'] - float(S.mse(x, y))) <= 1e-7 and abs(sc['psnr'] - float(S.psnr(x, y))) <= 1e-4
    assert abs(sc['ssim'] - float(S.ssim(x, y))) <= 2e-6
    paths = render.save_frames(frames.to(DEV), str(tmp_path / 'run'))
    assert [p.split('/')[-1] for p in paths] == ['img_000.png', 'img_001.png']
    want = (np.clip(frames.numpy(), 0, 1) * 255).astype(np.uint8)[..., ::-1]       # inference.py:260-262
    for p, w in zip(paths, want):
        assert np.array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), w)


@pytest.mark.parametrize('white,with_noise', [(1, False), (0, True)])
def test_raw2outputs_backward_matches_autograd_of_the_oracle(white, with_noise):
    """nrf_raw2outputs_backward against torch autograd through the oracle's composite (fp64 on the CPU): the
    gradient of a loss that reads rgb, weights and alpha, as the reference's MSE + GMM-density losses do."""
    from types import SimpleNamespace
    torch.manual_seed(5)
    B, n = 37, 192
    raw = torch.randn(B, n, 4) * torch.tensor([1., 1., 1., 3.])
    z = torch.sort(torch.rand(B, n) * 3 + 1, -1)[0]
    dirs = torch.randn(B, 1, 3).expand(B, n, 3).contiguous()
    noise = torch.randn(B, n) if with_noise else None
    wr, ww, wa = torch.randn(B, 3), torch.randn(B, n), torch.randn(B, n) * 0.1
    r64 = raw.double().requires_grad_(True)
    rgb, w, a = O.composite(r64, z.double(), dirs.double(), white_background=white, noise=None if noise is None else noise.double())
    ((rgb * wr.double()).sum() + (w * ww.double()).sum() + (a * wa.double()).sum()).backward()
    want = r64.grad

    args = SimpleNamespace(sigma_noise_std=0.0, white_background=white)
    rg = raw.to(DEV).requires_grad_(True)
    if noise is None:
        rgb_g, w_g, a_g = ops.raw2outputs(rg, z.to(DEV), dirs.to(DEV), args)
    else:      # the public wrapper draws its own noise; drive the autograd function with the test's draw
        rgb_g, w_g, a_g = ops._Raw2Outputs.apply(rg, z.to(DEV), dirs.to(DEV), noise.to(DEV), white)
    assert float((rgb_g.detach().cpu().double() - rgb.detach()).abs().max()) < 1e-5
    ((rgb_g * wr.to(DEV)).sum() + (w_g * ww.to(DEV)).sum() + (a_g * wa.to(DEV)).sum()).backward()
    got = rg.grad.cpu().double()
    scale = want.abs().max()
    assert float((got - want).abs().max()) <= 2e-5 * float(scale), float((got - want).abs().max() / scale)
    # no gradient requested -> plain forward, no graph
    out = ops.raw2outputs(raw.to(DEV), z.to(DEV), dirs.to(DEV), args)
    assert not out[0].requires_grad


@pytest.mark.parametrize('L,ident', [(10, False), (4, True)])
def test_positional_encoding_backward(L, ident):
    torch.manual_seed(6)
    x = torch.rand(50, 7, 3) * 2 - 1
    gout = torch.randn(50, 7, 3 * ((1 if ident else 0) + 2 * L))
    x64 = x.double().requires_grad_(True)
    enc = O.Encoder(L, ident)
    enc.bands = enc.bands.double()
    (enc.encode(x64) * gout.double()).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    (ops.PositionalEncoder(L, ident).encode(xg) * gout.to(DEV)).sum().backward()
    want = x64.grad
    assert float((xg.grad.cpu().double() - want).abs().max()) <= 2e-6 * float(want.abs().max())


@pytest.mark.parametrize('kind,n_layers,skips,width', [('nerf', 8, (4,), 256), ('append', 4, (2,), 128), ('smpl', 6, (3,), 256)])
def test_standalone_net_forward(kind, n_layers, skips, width):
    """RenderRayNet.forward / WarpFieldNet.forward on pre-encoded features (models/render_ray_net.py:42-61,
    models/warp_field_net.py:17-21): every nn.Linear is one tcgen05 GEMM launch; against the oracle's nets in fp64."""
    from smpl_nerf_b200.models import RenderRayNet, WarpFieldNet
    c, _, w, *_ = O.build_nets(kind, 3, 'default', n_layers=n_layers, skips=skips, width=width)
    net = RenderRayNet(n_layers, width, c.positions_dim, c.direcions_dim, c.additional_input_dim, list(skips))
    net.load_state_dict(c.state_dict())
    net = net.to(DEV).eval()
    torch.manual_seed(5)
    x = torch.randn(3, 333, c.positions_dim + c.additional_input_dim + c.direcions_dim)          # ragged row count, leading dims kept
    got = net(x.to(DEV))
    with torch.no_grad():
        want = c.double()(x.double())
    assert got.shape == want.shape == (3, 333, 4)
    assert float((got.double().cpu() - want).abs().max()) <= 2e-5 * (1 + float(want.abs().max()))
    with torch.no_grad():                                       # training-mode nets are fine under no_grad ...
        net.train()
        assert torch.equal(net(x.to(DEV)), got)
    with pytest.raises(NotImplementedError):                    # ... but the stand-alone forward does not build a graph
        net(x.to(DEV))
    with pytest.raises(RuntimeError):                           # and there is no CPU path
        net.eval()(x)
    if w is not None:
        wn = WarpFieldNet(n_layers, width, w.positions_dim, w.direcions_dim)
        wn.load_state_dict(w.state_dict())
        wn = wn.to(DEV).eval()
        xw = torch.randn(1000, w.positions_dim + w.direcions_dim)
        gw = wn(xw.to(DEV))
        assert gw.shape == (1000, 3)
        assert float((gw.double().cpu() - w.double()(xw.double())).abs().max()) <= 2e-5
