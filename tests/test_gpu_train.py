"""Training path: the tcgen05 GEMM building blocks against fp64 matmuls,
the layer-by-layer forward against the oracle, parameter gradients against fp64 autograd of the oracle, and the
solver's loop (solver/nerf_solver.py:76-88: MSE coarse + fine, Adam) against the same loop run with the oracle on the CPU."""
import copy
import ctypes as C

import pytest
import torch

from tests import helpers as H
from oracle import nerf_oracle as O
from smpl_nerf_b200 import _lib, engine, scene
from smpl_nerf_b200.models import (AppendSmplParamsPipeline, AppendToNerfPipeline, NerfPipeline, SmplNerfPipeline)

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _planes(x, pad=None, exact=False):
    """fp32 [rows, cols] -> (hi, lo[, ll]) planes through nrf_split_planes (fp16 pair, or the exact mode's bfloat16 triple)."""
    L = _lib.lib()
    rows, cols = x.shape
    pad = pad or cols
    dt = torch.bfloat16 if exact else torch.float16
    hi, lo = torch.empty(rows, pad, dtype=dt, device=DEV), torch.empty(rows, pad, dtype=dt, device=DEV)
    ll = torch.empty(rows, pad, dtype=dt, device=DEV) if exact else None
    _lib.check(L.nrf_split_planes(x.data_ptr(), rows, cols, cols, hi.data_ptr(), lo.data_ptr(), ll.data_ptr() if exact else None, pad, pad, None), 'split')
    return (hi, lo, ll) if exact else (hi, lo)


@pytest.mark.parametrize('S,K,N', [(1000, 256, 256), (300, 576, 512), (77, 64, 128)])
def test_gemm_exact_mode(S, K, N):
    """passes = 6 on bfloat16 hi/lo/ll planes: the split is exact (hi + lo + ll == x bit for bit), the product matches fp64 to
    fp32-accumulation accuracy, also for operands far outside the fp16 range."""
    L = _lib.lib()
    torch.manual_seed(S + K)
    a = torch.randn(S, K, device=DEV) * 3e5
    w = torch.randn(N, K, device=DEV) / K ** .5 * 1e-5
    ah, al, aq = _planes(a, exact=True)
    wh, wl, wq = _planes(w, exact=True)
    assert torch.equal(ah.float() + al.float() + aq.float(), a) and torch.equal(wh.float() + wl.float() + wq.float(), w)
    out = torch.empty(S, N, device=DEV)
    oh, ol, oq = (torch.empty(S, N, dtype=torch.bfloat16, device=DEV) for _ in range(3))
    _lib.check(L.nrf_gemm_planes(0, ah.data_ptr(), al.data_ptr(), aq.data_ptr(), S, K, wh.data_ptr(), wl.data_ptr(), wq.data_ptr(), N, 6, None, 0,
                                 out.data_ptr(), oh.data_ptr(), ol.data_ptr(), oq.data_ptr(), None), 'gemm exact')
    want = a.double() @ w.double().t()
    err = float((out.double() - want).abs().max())
    print(f'exact gemm K={K}: max abs err {err:.2e} on outputs of magnitude {float(want.abs().max()):.1f}')
    assert err <= 3e-6 * max(1.0, (K / 256) ** .5) * float(want.abs().max())
    assert torch.equal(oh.float() + ol.float() + oq.float(), out)
    torch.cuda.synchronize()


@pytest.mark.parametrize('passes', [3, 1])
@pytest.mark.parametrize('S,K,N', [(1000, 256, 256), (128, 64, 256), (777, 128, 128), (4096, 256, 64), (50, 256, 128),
                                   (500, 1408, 256), (300, 576, 512)])      # the last two: K too large for a resident weight slice
def test_gemm_forward_and_dx(S, K, N, passes):
    """nrf_gemm_planes: C = A B^T (weights K-major, the forward of nn.Linear, bias + ReLU fused) and C = A B (weights
    N-major: dX = dY W) against fp64 matmuls of the same fp32 inputs."""
    L = _lib.lib()
    torch.manual_seed(S + K + N)
    a = torch.randn(S, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** .5
    bias = torch.randn(N, device=DEV)
    tol = (2e-5 if passes == 3 else 2e-2) * max(1.0, (K / 256) ** .5)
    ah, al = _planes(a)
    wh, wl = _planes(w)
    # forward with bias and ReLU, planes + fp32 copy out
    out = torch.empty(S, N, device=DEV)
    oh = torch.empty(S, N, dtype=torch.float16, device=DEV)
    ol = torch.empty(S, N, dtype=torch.float16, device=DEV)
    _lib.check(L.nrf_gemm_planes(0, ah.data_ptr(), al.data_ptr(), None, S, K, wh.data_ptr(), wl.data_ptr(), None, N, passes, bias.data_ptr(), 1,
                                 out.data_ptr(), oh.data_ptr(), ol.data_ptr(), None, None), 'gemm fwd')
    want = torch.relu(a.double() @ w.double().t() + bias.double())
    assert float((out.double() - want).abs().max()) <= tol
    assert float(((oh.double() + ol.double()) - out.double()).abs().max()) <= 1e-5 * (1 + float(out.abs().max()))
    # dX-type product: B is [K, N] row-major (N-major operand), fp32 out, no epilogue
    w2 = torch.randn(K, N, device=DEV) / K ** .5
    w2h, w2l = _planes(w2)
    out2 = torch.empty(S, N, device=DEV)
    _lib.check(L.nrf_gemm_planes(1, ah.data_ptr(), al.data_ptr(), None, S, K, w2h.data_ptr(), w2l.data_ptr(), None, N, passes, None, 0,
                                 out2.data_ptr(), None, None, None, None), 'gemm dx')
    assert float((out2.double() - a.double() @ w2.double()).abs().max()) <= tol
    torch.cuda.synchronize()


@pytest.mark.parametrize('passes', [3, 1])
@pytest.mark.parametrize('S,M,N', [(1000, 256, 256), (64, 128, 64), (5000, 128, 256), (333, 256, 128), (20000, 256, 64)])
def test_gemm_dw(S, M, N, passes):
    """nrf_gemm_dw: dW = dY^T X with both operands MN-major straight from the row-major planes, split over the SMs."""
    L = _lib.lib()
    torch.manual_seed(S + M + N)
    dy = torch.randn(S, M, device=DEV)
    x = torch.randn(S, N, device=DEV)
    dh, dl = _planes(dy)
    xh, xl = _planes(x)
    max_split = 148
    partial = torch.empty(max_split * M * N + 2, device=DEV)
    out = torch.zeros(M, N, device=DEV)
    _lib.check(L.nrf_gemm_dw(dh.data_ptr(), dl.data_ptr(), M, xh.data_ptr(), xl.data_ptr(), N, S, passes, partial.data_ptr(), max_split,
                             out.data_ptr(), None), 'gemm dw')
    want = dy.double().t() @ x.double()
    scale = float(want.abs().max())
    assert float((out.double() - want).abs().max()) <= (3e-6 if passes == 3 else 3e-3) * scale
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ pipelines
def _build(kind, seed=5, n_layers=8, skips=(4,), variant='dense'):
    return O.build_nets(kind, seed, variant, n_layers=n_layers, skips=skips)


def _pipe(kind, nets, args):
    c, f, w, pe, de, he = nets
    if kind == 'nerf':
        return NerfPipeline(c, f, args, pe, de)
    if kind == 'append':
        return AppendToNerfPipeline(c, f, args, pe, de, he)
    if kind == 'append_full':
        return AppendSmplParamsPipeline(c, f, args, pe, de, he)
    return SmplNerfPipeline(c, f, w, args, pe, de, he)


def _loss(out, gt):
    return torch.mean((out[0] - gt) ** 2) + torch.mean((out[1] - gt) ** 2)        # solver/nerf_solver.py:48-51


def _rays(kind, h, w, nc, seed):
    rays = scene.make_rays(h, w, nc, seed=seed, with_colours=True, arm_angle_deg=25.0)
    if kind == 'append_full':         # all 69 pose parameters carry values, different per ray (shuffled training batches)
        g = torch.Generator().manual_seed(seed)
        rays['goal_pose'] = torch.rand(rays['goal_pose'].shape, generator=g) * 1.2 - 0.6
    return scene.data_list(rays, kind)


@pytest.mark.parametrize('kind', ['nerf', 'append', 'append_full', 'smpl'])
def test_train_forward_matches_oracle(kind):
    """Under grad mode with training-mode nets the pipelines run the layer-by-layer path: same outputs as the oracle
    (fine pass teacher-forced on the oracle's depths, like the stage-wise inference tests)."""
    nets = _build(kind)
    args = O.make_args()
    data = _rays(kind, 6, 7, 64, 3)
    with torch.no_grad():
        want = H.run_oracle(kind, nets, args, data)
    gnets, gdata = H.to_cuda(nets, data)
    for m in gnets[:3]:
        if m is not None:
            m.train()
    got = engine.render(kind, gnets[0], gnets[1], gnets[2], args, gnets[3], gnets[4], gnets[5], gdata, z_all_in=want['z_all'].to(DEV))
    torch.cuda.synchronize()
    assert got['rgb'].requires_grad and got['rgb_fine'].requires_grad
    assert float((got['rgb'].detach().cpu() - want['rgb']).abs().max()) <= H.TOL_RGB
    assert float((got['rgb_fine'].detach().cpu() - want['rgb_fine']).abs().max()) <= H.TOL_RGB
    mask = H.alpha_mask_well_conditioned(want['raw_fine'][..., 3])
    assert float((got['alpha_out'].cpu() - want['alpha_out']).abs()[mask].max()) <= H.TOL_ALPHA
    assert torch.equal(got['samples_out'].cpu(), want['samples_out'])
    if kind == 'smpl':
        assert float((got['warped_out'].cpu() - want['warped_out']).abs().max()) <= 1e-4
        assert float((got['warp_out'].cpu() - want['warp_out']).abs().max()) <= 1e-4
    assert int(got['status'].item()) == 0


@pytest.mark.parametrize('kind,run_fine', [('nerf', 1), ('smpl', 1), ('append', 0), ('smpl', 0)])
def test_training_with_density_noise_and_without_fine_pass(kind, run_fine):
    """args.sigma_noise_std = 1 is the reference's TRAINING default (config_parser.py:87; utils.py:172-174 draws N(0, std) per
    sample in both passes) and run_fine = 0 is a shipped ablation: same draws -> same loss and gradients as torch autograd of the oracle."""
    nets = O.build_nets(kind, 17, 'dense', n_layers=4, skips=(1,))
    args = O.make_args(number_fine_samples=32, sigma_noise_std=1.0, run_fine=run_fine)
    data = _rays(kind, 6, 6, 32, 9)
    B = data[0].shape[0]
    torch.manual_seed(5)
    n_c, n_f = torch.randn(B, 32), torch.randn(B, 64)
    ref = [copy.deepcopy(m) if m is not None else None for m in nets[:3]]
    o = H.run_oracle(kind, (ref[0], ref[1], ref[2]) + tuple(nets[3:]), args, data, noise_coarse=n_c, noise_fine=n_f if run_fine else None)
    _loss((o['rgb'], o['rgb_fine']), data[-1]).backward()
    gnets, gdata = H.to_cuda(nets, data)
    for m in gnets[:3]:
        if m is not None:
            m.train()
    out = engine.render(kind, gnets[0], gnets[1], gnets[2], args, gnets[3], gnets[4], gnets[5], gdata,
                        noise=(n_c.to(DEV), n_f.to(DEV) if run_fine else None), z_all_in=o['z_all'].to(DEV) if run_fine else None)
    loss = _loss((out['rgb'], out['rgb_fine']), gdata[-1])
    loss.backward()
    torch.cuda.synchronize()
    assert float((out['rgb'].detach().cpu() - o['rgb']).abs().max()) <= H.TOL_RGB
    assert float((out['rgb_fine'].detach().cpu() - o['rgb_fine']).abs().max()) <= H.TOL_RGB
    if not run_fine:
        assert out['rgb_fine'] is out['rgb'] and gnets[1].positions_pose_input.weight.grad is None      # the fine net is not evaluated
    for net, r in zip(gnets[:3], ref):
        if net is None or (net is gnets[1] and not run_fine):
            continue
        for (pn, p), (_, q) in zip(net.named_parameters(), r.named_parameters()):
            rel = float((p.grad.cpu() - q.grad).norm() / (q.grad.norm() + 1e-30))
            assert rel <= 2e-2, f'{pn}: {rel:.2e}'          # fp32 torch autograd is the comparator here (its own noise: ~1e-2 behind the warp chain)


def _grad_check(kind, precision, tol, variant='dense', n_layers=8, skips=(4,), floor_factor=6.0, width=256, shape=(8, 8, 32, 64)):
    nets = O.build_nets(kind, 7, variant, n_layers=n_layers, skips=skips, width=width)
    args = O.make_args(number_fine_samples=shape[3])
    data = _rays(kind, shape[0], shape[1], shape[2], 11)
    with torch.no_grad():
        z_all = H.run_oracle(kind, nets, args, data)['z_all']
    # fp64 autograd of the oracle on the SAME depths
    n64 = [copy.deepcopy(m).double() if m is not None else None for m in nets[:3]]
    d64 = [t.double() for t in data]
    o64 = H.run_oracle(kind, (n64[0], n64[1], n64[2]) + tuple(nets[3:]), args, d64, z_all_in=z_all)
    loss64 = _loss((o64['rgb'], o64['rgb_fine']), d64[-1])
    loss64.backward()
    # the reference's OWN fp32 autograd on the same depths: its deviation from fp64 is the noise floor (the encodings of the
    # fine points o + d z and of the warped points carry an fp32 rounding of the point, amplified 2^9-fold by the highest
    # frequency -- that alone moves the first layer's weight gradient by ~1e-3 relative)
    n32 = [copy.deepcopy(m) if m is not None else None for m in nets[:3]]
    o32 = H.run_oracle(kind, (n32[0], n32[1], n32[2]) + tuple(nets[3:]), args, data, z_all_in=z_all)
    _loss((o32['rgb'], o32['rgb_fine']), data[-1]).backward()
    gnets, gdata = H.to_cuda(nets, data)
    for m in gnets[:3]:
        if m is not None:
            m.train()
    out = engine.render(kind, gnets[0], gnets[1], gnets[2], args, gnets[3], gnets[4], gnets[5], gdata, z_all_in=z_all.to(DEV), precision=precision)
    loss = _loss((out['rgb'], out['rgb_fine']), gdata[-1])
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(loss64)) <= 1e-5
    worst = 0.0
    for name, net, ref, r32 in zip(('coarse', 'fine', 'warp'), gnets[:3], n64, n32):
        if net is None:
            continue
        for (pn, p), (_, q), (_, q32) in zip(net.named_parameters(), ref.named_parameters(), r32.named_parameters()):
            assert p.grad is not None, f'{name}.{pn} received no gradient'
            g, w = p.grad.double().cpu(), q.grad
            rel = float((g - w).norm() / (w.norm() + 1e-30))
            floor = float((q32.grad.double() - w).norm() / (w.norm() + 1e-30))
            worst = max(worst, rel)
            assert rel <= tol + floor_factor * floor, \
                f'{name}.{pn}: relative gradient error {rel:.2e} (reference fp32-vs-fp64: {floor:.2e}; |g| = {float(w.norm()):.3e})'
    return worst


@pytest.mark.parametrize('kind', ['nerf', 'append', 'append_full', 'smpl'])
def test_parameter_gradients_match_fp64_autograd(kind):
    """VERDICT r1 'done' criterion: every parameter gradient of the solver's loss within 1e-3 (relative, per tensor) of fp64
    autograd of the oracle, plus 6x the deviation the reference's OWN fp32 autograd shows on that tensor: the engine's hi/lo
    operands carry 22 significant bits against fp32's 24, so cancellation-dominated gradients (tiny |g|: the sigma head, the
    2^9-frequency columns of the first layer, everything behind the SMPL warp chain, where the reference itself is only
    good to 1e-2) see up to ~4x the reference's fp32 rounding noise.  tools/dbg_grad.py lists both per tensor."""
    worst = _grad_check(kind, 0, 1e-3)
    print(f'{kind}: worst relative gradient error {worst:.2e}')


@pytest.mark.parametrize('kind', ['nerf', 'smpl'])
def test_parameter_gradients_one_pass(kind):
    """precision = 1 (one fp16 MMA pass forward and backward, mixed-precision training): gradients to a few 1e-2."""
    worst = _grad_check(kind, 1, 1e-1)
    print(f'{kind} (1 pass): worst relative gradient error {worst:.2e}')


@pytest.mark.parametrize('kind', ['nerf', 'smpl'])
def test_gradients_ragged_sample_counts(kind):
    """Sample counts that are no multiple of the 128 / 256-row GEMM tiles or of the 64-sample dW chunks (35 rays x 32 coarse = 1,120
    rows, x 61 = 2,135 rows in the fine pass): TMA zero-fill of the ragged last tile, clipped TMA stores, bit-mask rows, per-ray bias.
    (With only 24 coarse samples the SMPL coarse net's first-layer gradient is 5e-6 in norm and the 22-bit planes show 2e-3 on it at
    ragged and aligned counts alike -- tools/dbg_ragged.py -- so the case keeps 32.)"""
    _grad_check(kind, 0, 1e-3, shape=(5, 7, 32, 29))


def test_gradients_shallow_net_no_skip_sharp_weights():
    """A different architecture (depth 4, no skip: BASELINE configs[0]) and the ill-conditioned 'sharp' weights (x2)."""
    _grad_check('nerf', 0, 1e-3, variant='sharp', n_layers=4, skips=())


@pytest.mark.parametrize('kind,width', [('nerf', 128), ('smpl', 128), ('append', 512), ('smpl', 512)])
def test_other_hidden_widths(kind, width):
    """netwidth / netwidth_fine / netwidth_warp are flags in the reference (config_parser.py:20,24,30).  Widths 128 and 512 run
    on the layer-by-layer path -- for inference (eval mode, no_grad) as well as for training; 256 keeps the fused kernel."""
    nets = O.build_nets(kind, 9, 'dense', n_layers=5, skips=(2,), width=width)
    args = O.make_args(number_fine_samples=64)
    data = _rays(kind, 5, 6, 32, 4)
    with torch.no_grad():
        want = H.run_oracle(kind, nets, args, data)
    gnets, gdata = H.to_cuda(nets, data)
    for m in gnets[:3]:
        if m is not None:
            m.eval()
    with torch.no_grad():
        got = engine.render(kind, gnets[0], gnets[1], gnets[2], args, gnets[3], gnets[4], gnets[5], gdata, z_all_in=want['z_all'].to(DEV))
    torch.cuda.synchronize()
    assert not got['rgb'].requires_grad
    assert float((got['rgb'].cpu() - want['rgb']).abs().max()) <= H.TOL_RGB
    assert float((got['rgb_fine'].cpu() - want['rgb_fine']).abs().max()) <= H.TOL_RGB
    mask = H.alpha_mask_well_conditioned(want['raw_fine'][..., 3])
    assert float((got['alpha_out'].cpu() - want['alpha_out']).abs()[mask].max()) <= H.TOL_ALPHA
    assert int(got['status'].item()) == 0
    worst = _grad_check(kind, 0, 1e-3, n_layers=5, skips=(2,), width=width)
    print(f'{kind} width {width}: worst relative gradient error {worst:.2e}')


@pytest.mark.parametrize('request_scale', [1e5, 1e7])      # activations beyond 65504; at 1e7 the first layer's WEIGHTS too
def test_exact_mode_beyond_the_fp16_range(request_scale):
    """precision = 2 (bf16 x 3 planes, six MMA passes, layer by layer): operands are exact fp32 values with fp32's exponent
    range.  A net whose hidden activations reach 2e5 / 2e7 (first layer x scale, second layer / scale: the same function) trips the
    fused kernel's fp16 range flag; the exact mode renders it within the parity bars."""
    scale = request_scale
    nets = O.build_nets('nerf', 13, 'dense')
    with torch.no_grad():
        for net in nets[:2]:
            net.positions_pose_input.weight.mul_(scale); net.positions_pose_input.bias.mul_(scale)
            net.positional_net[0].weight.div_(scale)
    args = O.make_args()
    data = _rays('nerf', 6, 6, 64, 2)
    with torch.no_grad():
        want = H.run_oracle('nerf', nets, args, data)
    gnets, gdata = H.to_cuda(nets, data)
    with torch.no_grad():
        par = engine.render('nerf', gnets[0], gnets[1], None, args, gnets[3], gnets[4], None, gdata, z_all_in=want['z_all'].to(DEV))
        got = engine.render('nerf', gnets[0], gnets[1], None, args, gnets[3], gnets[4], None, gdata, taps=True,
                            z_all_in=want['z_all'].to(DEV), precision=2)
    torch.cuda.synchronize()
    assert int(par['status'].item()) & 1, 'the fp16 range flag of the fused kernel did not fire'
    assert int(got['status'].item()) == 0
    assert float((got['raw_coarse'][..., 3].cpu() - want['raw_coarse'][..., 3]).abs().max()) <= H.TOL_SIGMA * 3     # activations ~1e6: the reference's own fp32 noise is of this order
    assert float((got['rgb_fine'].cpu() - want['rgb_fine']).abs().max()) <= H.TOL_RGB
    mask = H.alpha_mask_well_conditioned(want['raw_fine'][..., 3])
    assert float((got['alpha_out'].cpu() - want['alpha_out']).abs()[mask].max()) <= H.TOL_ALPHA


def test_exact_mode_meets_the_literal_sigma_bar_on_trained_weights():
    """on the trained vanilla-NeRF checkpoint (hidden activations O(100), sigma up to 240) the parity
    mode's 22-bit operands miss the raw-sigma reading of the 1e-4 bar by up to 7x; the exact mode meets it literally."""
    ck, nets, args = H.load_trained()
    with torch.no_grad():
        want = H.run_oracle('nerf', nets, args, ck['data'])
    gnets, gdata = H.to_cuda(nets, ck['data'])
    with torch.no_grad():
        got = engine.render('nerf', gnets[0], gnets[1], None, args, gnets[3], gnets[4], None, gdata, taps=True,
                            z_all_in=want['z_all'].to(DEV), precision=2)
    torch.cuda.synchronize()
    for k in ('raw_coarse', 'raw_fine'):
        err = float((got[k][..., 3].cpu() - want[k][..., 3]).abs().max())
        print(f'exact mode, trained d4 checkpoint: max |sigma - reference| {k} = {err:.2e} (max |sigma| {float(want[k][..., 3].abs().max()):.0f})')
        assert err <= H.TOL_SIGMA
    assert float((got['rgb_fine'].cpu() - want['rgb_fine']).abs().max()) <= 1e-5


def test_eval_mode_and_no_grad_stay_on_the_fused_kernel():
    """inference.py:247-254 calls the pipeline with eval-mode nets and autograd on; validation uses torch.no_grad():
    both must take the fused inference kernel (graph-less outputs), training-mode nets the differentiable path."""
    nets = _build('nerf')
    args = O.make_args()
    gnets, gdata = H.to_cuda(nets, _rays('nerf', 4, 4, 64, 1))
    pipe = _pipe('nerf', gnets, args)
    gnets[0].eval(); gnets[1].eval()
    out = pipe(gdata)
    assert not out[0].requires_grad and not out[1].requires_grad
    gnets[0].train(); gnets[1].train()
    with torch.no_grad():
        out = pipe(gdata)
    assert not out[0].requires_grad
    out = pipe(gdata)
    assert out[0].requires_grad and out[1].requires_grad and not out[2].requires_grad and not out[3].requires_grad
    with pytest.raises(ValueError):
        engine.render('nerf', gnets[0], gnets[1], None, args, gnets[3], gnets[4], None, gdata, trace_cap=16)


def test_gradients_are_linear_in_the_batch_at_full_size():
    """Size-independent property at the training workload's full size (2,048 rays, 64 + 128 samples, 8x256 nets + warp net, where the
    CPU oracle would take minutes): the loss is a mean over rays, so the gradient of the whole batch equals the mean of the gradients
    of its two halves (different tile counts, different per-layer gradient scales, different dW split counts)."""
    nets = _build('smpl', 11)
    args = O.make_args(number_fine_samples=128)
    rays = scene.make_rays(64, 32, 64, seed=3, with_colours=True, arm_angle_deg=40.0)
    data = scene.data_list(rays, 'smpl')
    gnets, gdata = H.to_cuda(nets, data)
    for m in gnets[:3]:
        m.train()
    pipe = _pipe('smpl', gnets, args)
    params = [p for m in gnets[:3] for p in m.parameters()]
    names = [f'{k}.{n}' for k, m in zip(('coarse', 'fine', 'warp'), gnets[:3]) for n, _ in m.named_parameters()]

    def grads(sl):
        for p in params:
            p.grad = None
        d = [t[sl] for t in gdata]
        _loss(pipe(d), d[-1]).backward()
        return [p.grad.double() for p in params]

    full, a, b = grads(slice(0, 2048)), grads(slice(0, 1024)), grads(slice(1024, 2048))
    for n, f, x, y in zip(names, full, a, b):
        want = 0.5 * (x + y)
        err = float((f - want).norm() / (want.norm() + 1e-30))
        assert err <= 2e-4, f'{n}: |grad(batch) - mean(grad(halves))| / |.| = {err:.2e}'


def test_workspace_cache_and_two_live_graphs():
    """The training workspace is cached per device between steps.  A second forward issued while the first graph is still alive
    must not share it; a backward may run only once; a graph that is dropped without a backward hands the buffer back."""
    from smpl_nerf_b200 import train as TR
    nets = _build('nerf', 3, 4, (2,))
    args = O.make_args(number_fine_samples=32)
    d1 = _rays('nerf', 6, 6, 32, 2)
    d2 = _rays('nerf', 6, 6, 32, 9)
    gnets, g1 = H.to_cuda(nets, d1)
    _, g2 = H.to_cuda(nets, d2)
    for m in gnets[:2]:
        m.train()
    pipe = _pipe('nerf', gnets, args)
    params = [p for m in gnets[:2] for p in m.parameters()]

    def grads(data):
        for p in params:
            p.grad = None
        _loss(pipe(data), data[-1]).backward()
        return [p.grad.clone() for p in params]

    ref1, ref2 = grads(g1), grads(g2)
    dev = g1[0].device.index
    grads(g1)
    torch.cuda.synchronize()
    m0 = torch.cuda.memory_allocated()
    for _ in range(3):                                        # no per-step leak (outputs, workspace) once the graphs are gone
        grads(g1)
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() <= m0
    assert TR._ws_cache[dev][1] is False                      # released by the backward
    cached = TR._ws_cache[dev][0]
    # two graphs alive at once: the second forward gets a private buffer, both backward passes reproduce the separate runs
    for p in params:
        p.grad = None
    o1 = pipe(g1)
    assert TR._ws_cache[dev][1] is True and TR._ws_cache[dev][0] is cached
    o2 = pipe(g2)
    l1, l2 = _loss(o1, g1[-1]), _loss(o2, g2[-1])
    l2.backward()
    got2 = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    l1.backward()
    got1 = [p.grad.clone() for p in params]
    for a_, b_ in zip(got1 + got2, ref1 + ref2):
        assert torch.allclose(a_, b_, rtol=1e-4, atol=1e-7 * float(b_.abs().max()) + 1e-12)      # (head gradients use float atomics)
    assert TR._ws_cache[dev][1] is False
    with pytest.raises(RuntimeError):
        l1.backward()
    # a graph that is never back-propagated frees the cached buffer when it dies
    o3 = pipe(g1)
    assert TR._ws_cache[dev][1] is True
    del o3
    import gc
    gc.collect()
    assert TR._ws_cache[dev][1] is False
    TR.release_workspaces()
    assert dev not in TR._ws_cache


@pytest.mark.parametrize('kind', ['nerf', 'append', 'smpl'])
def test_solver_loop_tracks_the_reference_loop(kind):
    """solver/nerf_solver.py:76-88 / solver/smpl_nerf_solver.py:66-83: 50 Adam steps on the drop-in pipeline (GPU) and on
    the oracle port of the reference pipeline (CPU, torch autograd) from the same initial weights and the same batches.
    Both must learn, start identically, and end at the same loss level.  (Adam divides every gradient by its running
    magnitude, so parameters whose gradient is rounding noise take full-size steps in a noise-given direction: the two
    trajectories separate during the first, fast phase of training -- most for the SMPL pipeline, whose warp-net gradients are
    only good to 1e-2 in fp32 on EITHER side -- and meet again; pointwise the curves are compared loosely, the level tightly.)"""
    torch.manual_seed(0)
    nets = _build(kind, 21, 4, (2,), 'dense')
    args = O.make_args(number_fine_samples=32, sigma_noise_std=0.)
    rays = scene.make_rays(24, 24, 32, seed=5, with_colours=True, arm_angle_deg=35.0)
    n = rays['z_vals'].shape[0]
    fg = torch.nonzero((rays['rgb'] < 0.99).any(-1)).flatten()
    gnets, _ = H.to_cuda(nets, [])
    cpu = [copy.deepcopy(m) if m is not None else None for m in nets[:3]]
    for m in list(gnets[:3]) + cpu:
        if m is not None:
            m.train()
    pipe = _pipe(kind, gnets, args)
    opt_g = torch.optim.Adam([p for m in gnets[:3] if m is not None for p in m.parameters()], lr=5e-4)
    opt_c = torch.optim.Adam([p for m in cpu if m is not None for p in m.parameters()], lr=5e-4)
    g = torch.Generator().manual_seed(1)
    lg, lc = [], []
    for step in range(50):
        sel = torch.cat([torch.randint(0, n, (96,), generator=g), fg[torch.randint(0, fg.numel(), (64,), generator=g)]])
        data = scene.data_list(rays, kind, sel)
        out = pipe([t.to(DEV) for t in data])
        loss = _loss(out, data[-1].to(DEV))
        opt_g.zero_grad(); loss.backward(); opt_g.step()
        lg.append(float(loss.detach()))
        o = H.run_oracle(kind, (cpu[0], cpu[1], cpu[2]) + tuple(nets[3:]), args, data)
        loss_c = _loss((o['rgb'], o['rgb_fine']), data[-1])
        opt_c.zero_grad(); loss_c.backward(); opt_c.step()
        lc.append(float(loss_c.detach()))
    lg, lc = torch.tensor(lg), torch.tensor(lc)
    print(f'{kind}: reference {[round(float(x), 4) for x in lc[::5]]}')
    print(f'{kind}: engine    {[round(float(x), 4) for x in lg[::5]]}')
    assert float(lc[-10:].mean()) < 0.8 * float(lc[:3].mean()), 'the reference loop itself did not learn'
    assert float(lg[-10:].mean()) < 0.8 * float(lg[:3].mean()), 'the engine loop did not learn'
    # after ONE Adam step (every parameter moves by lr * sign(g)): nerf / append agree to 2e-3; for smpl the reference's own fp32
    # and fp64 warp-net gradients differ by 10-30 % in this free-running configuration (tools/dbg_grad.py smpl loop), so do the steps
    assert abs(float(lg[0]) - float(lc[0])) <= 1e-4 and abs(float(lg[1]) - float(lc[1])) <= (0.1 if kind == 'smpl' else 2e-3) * float(lc[1])
    assert float((lg - lc).abs().max()) <= (0.3 if kind == 'smpl' else 0.1) * float(lc.max())
    assert abs(float(lg[-10:].mean()) - float(lc[-10:].mean())) <= 0.15 * float(lc[-10:].mean())
