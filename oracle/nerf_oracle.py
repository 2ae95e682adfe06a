"""torch-CPU restatement of the SMPL-NeRF pipeline forward -- TEST INFRASTRUCTURE ONLY.

This is the parity oracle for the CUDA engine.  It restates, op for op (so that fp32
results are bit-identical to the reference on the same torch build), what the reference
computes in

  * utils.py:114-131     PositionalEncoder           -> ``Encoder``
  * models/render_ray_net.py:6-61   RenderRayNet     -> ``RayNet``
  * models/warp_field_net.py:6-21   WarpFieldNet     -> ``WarpNet``
  * utils.py:134-191     raw2outputs                 -> ``composite``
  * utils.py:194-228     sample_pdf                  -> ``inverse_cdf``
  * utils.py:231-264     fine_sampling               -> ``fine_samples``
  * models/nerf_pipeline.py:14-67            -> ``nerf_forward``
  * models/smpl_nerf_pipeline.py:16-100      -> ``smpl_nerf_forward``
  * models/append_to_nerf_pipeline.py:14-90  -> ``append_to_nerf_forward``

The reference's ``torchsearchsorted.searchsorted(cdf, u, side='right')`` is replaced by
``torch.searchsorted(cdf, u, right=True)`` (index-identical; the C
restatement of the bisection itself is oracle/searchsorted_oracle.c).

Parity status: PINNED -- tests/test_oracle_vs_reference.py checks every function here
bit-for-bit against the imported reference where /root/reference exists, and
tests/test_golden.py checks it against fixtures minted from the reference
(tests/golden/make_golden.py) everywhere else.

All functions work in the dtype of their inputs (fp32 = the reference; fp64 = the
"noise floor" run used to put parity errors in context).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------
# a1: positional encoding (utils.py:114-131)
# --------------------------------------------------------------------------------------
class Encoder:
    """sin/cos features: [x?] ++ [sin(2^k x), cos(2^k x)] for k < L, frequency-major.

    Same attribute names as the reference class (``number_frequencies``,
    ``include_identity``, ``output_dim`` = features per scalar) because the drop-in
    pipelines read them.
    """

    def __init__(self, number_frequencies: int, include_identity):
        self.number_frequencies = int(number_frequencies)
        self.include_identity = include_identity
        # the reference builds its bands with pow(2, linspace) -- exactly 2^k in fp32
        self.bands = torch.pow(2, torch.linspace(0., number_frequencies - 1, number_frequencies)) \
            if number_frequencies > 0 else torch.zeros(0)
        self.output_dim = (1 if include_identity else 0) + 2 * self.number_frequencies

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        parts = [x] if self.include_identity else []
        for band in self.bands:           # 0-dim fp32 tensor, like the reference's ``freq``
            scaled = x * band
            parts.append(torch.sin(scaled))
            parts.append(torch.cos(scaled))
        return torch.cat(parts, -1)


# --------------------------------------------------------------------------------------
# a2 / a3: the MLPs.  Parameter names and construction ORDER match the reference so that
# (i) state_dicts are interchangeable and (ii) torch.manual_seed(s) followed by
# construction yields the same default-initialised weights as the reference classes.
# --------------------------------------------------------------------------------------
class RayNet(nn.Module):
    def __init__(self, n_layers=8, width=256, positions_dim=60, directions_dim=24,
                 additional_input_dim=0, skips=(4,), use_directional_input=1):
        super().__init__()
        self.n_layers, self.width = n_layers, width
        self.positions_dim, self.direcions_dim = positions_dim, directions_dim  # (sic) reference spelling
        self.skips = list(skips)
        self.additional_input_dim = additional_input_dim
        self.use_directional_input = use_directional_input
        in0 = positions_dim + additional_input_dim
        self.positions_pose_input = nn.Linear(in0, width)
        self.positional_net = nn.ModuleList(
            [nn.Linear(width + in0 if i in self.skips else width, width) for i in range(n_layers - 1)])
        self.additional_linear_layer = nn.Linear(width, width)
        self.sigma_out_layer = nn.Linear(width, 1)
        half = width // 2
        self.directional_input = nn.Linear(width + directions_dim if use_directional_input else width, half)
        self.directional_net = nn.ModuleList([nn.Linear(half, half)])
        self.rgb_out_layer = nn.Linear(half, 3)

    def forward(self, x: torch.Tensor, tap: Optional[dict] = None) -> torch.Tensor:
        n_pp = self.positions_dim + self.additional_input_dim
        pp, dirs = x[..., :n_pp], x[..., -self.direcions_dim:]
        h = torch.relu(self.positions_pose_input(pp))
        for i, lin in enumerate(self.positional_net):
            h = torch.relu(lin(torch.cat([h, pp], -1) if i in self.skips else h))
        h = self.additional_linear_layer(h)
        sigma = self.sigma_out_layer(h)
        h = self.directional_input(torch.cat([h, dirs], -1) if self.use_directional_input else h)
        for lin in self.directional_net:
            h = torch.relu(lin(h))
        rgb = self.rgb_out_layer(h)
        return torch.cat([rgb, sigma], -1)


class WarpNet(nn.Module):
    def __init__(self, n_layers=8, width=256, positions_dim=60, pose_dim=24):
        super().__init__()
        self.positions_dim, self.direcions_dim = positions_dim, pose_dim
        self.linear1 = nn.Linear(positions_dim + pose_dim, width)
        self.linear2 = nn.Linear(width, 3)

    def forward(self, x):
        return self.linear2(torch.relu(self.linear1(x)))


# --------------------------------------------------------------------------------------
# a4: compositing (utils.py:134-191)
# --------------------------------------------------------------------------------------
def composite(raw: torch.Tensor, z: torch.Tensor, dirs: torch.Tensor, *, white_background,
              noise: Optional[torch.Tensor] = None):
    """raw[B,n,4], z[B,n], dirs[B,n,3] -> rgb[B,3], weights[B,n], alpha[B,n].

    ``noise`` (already scaled by sigma_noise_std) is added to the raw density before the
    ReLU; the reference draws it inside raw2outputs (utils.py:172-174) -- the oracle takes
    it as an argument so both sides of a parity test consume the same draw.
    """
    delta = z[..., 1:] - z[..., :-1]
    far = torch.tensor([1e10], dtype=z.dtype, device=z.device).expand(delta[..., :1].shape)
    delta = torch.cat([delta, far], -1)
    delta = delta * torch.norm(dirs, dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    sig = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-torch.relu(sig) * delta)
    keep = 1. - alpha + 1e-10
    ones = torch.ones(keep.shape[:-1], dtype=z.dtype, device=z.device).unsqueeze(-1)
    trans = torch.cumprod(torch.cat([ones, keep[..., :-1]], -1), -1)       # exclusive product
    weights = alpha * trans
    rgb = torch.sum(weights[..., None] * colour, -2)
    acc = torch.sum(weights, -1)
    if white_background:
        rgb = rgb + (1. - acc[..., None])
    return rgb, weights, alpha


# --------------------------------------------------------------------------------------
# a5 / a6: hierarchical sampling (utils.py:194-264)
# --------------------------------------------------------------------------------------
def inverse_cdf(bins: torch.Tensor, weights: torch.Tensor, n_fine: int) -> torch.Tensor:
    """bins[B,m], weights[B,m-1] -> deterministic inverse-CDF samples [B,n_fine]."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0., 1., steps=n_fine, dtype=torch.float32).to(device=cdf.device, dtype=cdf.dtype)
    u = u.expand(list(cdf.shape[:-1]) + [n_fine]).contiguous()
    idx = torch.searchsorted(cdf, u, right=True)           # == torchsearchsorted side='right'
    lo = torch.clamp(idx - 1, min=0)
    hi = torch.clamp(idx, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b0, b1 = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - c0) / denom
    return b0 + t * (b1 - b0)


def fine_samples(origin: torch.Tensor, direction: torch.Tensor, z: torch.Tensor,
                 weights: torch.Tensor, n_fine: int, z_all_in: Optional[torch.Tensor] = None):
    """``z_all_in`` (not in the reference): use these merged depths instead of the sampler's -- "teacher forcing" for
    stage-wise parity and gradient tests, where a flipped sampler decision would otherwise mask everything else."""
    mid = .5 * (z[..., 1:] + z[..., :-1])
    z_new = inverse_cdf(mid, weights[..., 1:-1], n_fine).detach()
    z_all, _ = torch.sort(torch.cat([z, z_new], -1), -1)
    if z_all_in is not None:
        z_all = z_all_in.to(z.dtype)
    pts = origin[..., None, :] + direction[..., None, :] * z_all[..., :, None]
    return z_all, pts, z_new


# --------------------------------------------------------------------------------------
# a8-a10: the three pipeline forwards
# --------------------------------------------------------------------------------------
def make_args(**kw) -> SimpleNamespace:
    """The six ``args`` fields the hot path reads."""
    d = dict(default_device=torch.device('cpu'), sigma_noise_std=0., white_background=1,
             run_fine=1, number_fine_samples=128, human_pose_encoding=1)
    d.update(kw)
    return SimpleNamespace(**d)


def _unit(v: torch.Tensor) -> torch.Tensor:
    return v / torch.norm(v, dim=-1, keepdim=True)


def _pose2(goal_pose: torch.Tensor) -> torch.Tensor:
    # the two arm angles the synthetic datasets vary (smpl_nerf_pipeline.py:28)
    return torch.stack([goal_pose[:, 38], goal_pose[:, 41]], -1)


def nerf_forward(coarse: RayNet, fine: RayNet, pos_enc: Encoder, dir_enc: Encoder, args,
                 data: Sequence[torch.Tensor], noise_coarse=None, noise_fine=None, z_all_in=None) -> Dict[str, torch.Tensor]:
    """models/nerf_pipeline.py:14-67.  Returns a dict; ``as_tuple`` gives the reference order."""
    samples, origin, direction, z = data[0], data[1], data[2], data[3]
    B, n = samples.shape[0], samples.shape[1]
    enc_x = pos_enc.encode(samples)
    dirs = direction[..., None, :].expand(B, n, 3)
    enc_d = dir_enc.encode(_unit(dirs))
    raw = coarse(torch.cat([enc_x.view(-1, enc_x.shape[-1]), enc_d.view(-1, enc_d.shape[-1])], -1)).view(B, n, 4)
    rgb, w, alpha = composite(raw, z, dirs, white_background=args.white_background, noise=noise_coarse)
    out = dict(rgb=rgb, raw_coarse=raw, weights_coarse=w, alpha_coarse=alpha, kind='nerf')
    if not args.run_fine:
        out.update(rgb_fine=rgb, samples_out=samples, alpha_out=alpha)
        return out
    z_all, pts, z_new = fine_samples(origin, direction, z, w, args.number_fine_samples, z_all_in)
    m = pts.shape[1]
    enc_xf = pos_enc.encode(pts)
    enc_df = enc_d[..., :1, :].expand(B, m, enc_d.shape[-1])
    raw_f = fine(torch.cat([enc_xf.view(-1, enc_xf.shape[-1]), enc_df.reshape(-1, enc_df.shape[-1])], -1)).reshape(B, m, 4)
    dirs_f = direction[..., None, :].expand(B, m, 3)
    rgb_f, w_f, alpha_f = composite(raw_f, z_all, dirs_f, white_background=args.white_background, noise=noise_fine)
    out.update(rgb_fine=rgb_f, samples_out=pts, alpha_out=alpha_f, z_new=z_new, z_all=z_all,
               raw_fine=raw_f, weights_fine=w_f)
    return out


def append_to_nerf_forward(coarse: RayNet, fine: RayNet, pos_enc: Encoder, dir_enc: Encoder,
                           pose_enc: Encoder, args, data, noise_coarse=None, noise_fine=None, full_pose=False, z_all_in=None):
    """models/append_to_nerf_pipeline.py:14-90 (pose features FIRST in the MLP input).
    ``full_pose``: models/append_smpl_params_pipeline.py:14-91 -- the same forward with all 69 pose
    parameters (encoded: 69 * 2L = 1380 features) instead of the two arm angles."""
    samples, origin, direction, z, goal_pose = data[0], data[1], data[2], data[3], data[4]
    B, n = samples.shape[0], samples.shape[1]
    pose = goal_pose if full_pose else _pose2(goal_pose)
    pose_feat = pose_enc.encode(pose) if args.human_pose_encoding else pose
    enc_x = pos_enc.encode(samples)
    dirs = direction[..., None, :].expand(B, n, 3)
    enc_d = dir_enc.encode(_unit(dirs))

    def run(net, enc_pts, rows):
        pf = pose_feat[..., None, :].expand(B, rows, pose_feat.shape[-1])
        ed = enc_d[..., :1, :].expand(B, rows, enc_d.shape[-1])
        x = torch.cat([pf.reshape(-1, pf.shape[-1]), enc_pts.view(-1, enc_pts.shape[-1]),
                       ed.reshape(-1, ed.shape[-1])], -1)
        return net(x).reshape(B, rows, 4)

    raw = run(coarse, enc_x, n)
    rgb, w, alpha = composite(raw, z, dirs, white_background=args.white_background, noise=noise_coarse)
    out = dict(rgb=rgb, raw_coarse=raw, weights_coarse=w, alpha_coarse=alpha, kind='append_full' if full_pose else 'append')
    if not args.run_fine:
        out.update(rgb_fine=rgb, samples_out=samples, alpha_out=alpha)
        return out
    z_all, pts, z_new = fine_samples(origin, direction, z, w, args.number_fine_samples, z_all_in)
    m = pts.shape[1]
    raw_f = run(fine, pos_enc.encode(pts), m)
    dirs_f = direction[..., None, :].expand(B, m, 3)
    rgb_f, w_f, alpha_f = composite(raw_f, z_all, dirs_f, white_background=args.white_background, noise=noise_fine)
    out.update(rgb_fine=rgb_f, samples_out=pts, alpha_out=alpha_f, z_new=z_new, z_all=z_all,
               raw_fine=raw_f, weights_fine=w_f)
    return out


def append_smpl_params_forward(coarse, fine, pos_enc, dir_enc, pose_enc, args, data, noise_coarse=None, noise_fine=None, z_all_in=None):
    """models/append_smpl_params_pipeline.py:14-91."""
    return append_to_nerf_forward(coarse, fine, pos_enc, dir_enc, pose_enc, args, data, noise_coarse, noise_fine,
                                  full_pose=True, z_all_in=z_all_in)


def smpl_nerf_forward(coarse: RayNet, fine: RayNet, warp: WarpNet, pos_enc: Encoder, dir_enc: Encoder,
                      pose_enc: Encoder, args, data, noise_coarse=None, noise_fine=None, z_all_in=None):
    """models/smpl_nerf_pipeline.py:16-100 (warp field, per-sample view directions)."""
    samples, origin, direction, z, goal_pose = data[0], data[1], data[2], data[3], data[4]
    B, n = samples.shape[0], samples.shape[1]
    pose = _pose2(goal_pose)
    pose_feat = pose_enc.encode(pose)

    def warp_of(pts, rows, encoded=True):
        if encoded:
            e = pos_enc.encode(pts)
            pf = pose_feat[..., None, :].expand(B, rows, pose_feat.shape[-1])
            x = torch.cat([e.reshape(-1, e.shape[-1]), pf.reshape(-1, pf.shape[-1])], -1)
        else:   # human_pose_encoding=0: raw xyz + raw pose (coarse pass only, reference :41-45)
            pf = pose[..., None, :].expand(B, rows, 2)
            x = torch.cat([pts.reshape(-1, 3), pf.reshape(-1, 2)], -1)
        return warp(x).view(pts.shape)

    def render(net, warped, rows):
        view = warped - origin[:, None, :]
        e = pos_enc.encode(warped)
        d = dir_enc.encode(_unit(view))
        x = torch.cat([e.view(-1, e.shape[-1]), d.view(-1, d.shape[-1])], -1)
        return net(x).view(B, rows, 4), view

    wf = warp_of(samples, n, encoded=bool(args.human_pose_encoding))
    warped = samples + wf
    raw, view = render(coarse, warped, n)
    # NB: coarse deltas are scaled by |warped - o| per sample (reference quirk, :52,:63)
    rgb, w, alpha = composite(raw, z, view, white_background=args.white_background, noise=noise_coarse)
    out = dict(rgb=rgb, raw_coarse=raw, weights_coarse=w, alpha_coarse=alpha, kind='smpl')
    if not args.run_fine:
        out.update(rgb_fine=rgb, warp_out=wf, samples_out=samples, warped_out=warped, alpha_out=alpha)
        return out
    z_all, pts, z_new = fine_samples(origin, direction, z, w, args.number_fine_samples, z_all_in)
    m = pts.shape[1]
    wf_f = warp_of(pts, m, encoded=True)                    # always the encoded form (:71-77)
    warped_f = pts + wf_f
    raw_f, _ = render(fine, warped_f, m)
    dirs_f = direction[..., None, :].expand(B, m, 3)        # fine deltas use |ray_direction| (:95-98)
    rgb_f, w_f, alpha_f = composite(raw_f, z_all, dirs_f, white_background=args.white_background, noise=noise_fine)
    out.update(rgb_fine=rgb_f, warp_out=wf_f, samples_out=pts, warped_out=warped_f, alpha_out=alpha_f,
               z_new=z_new, z_all=z_all, raw_fine=raw_f, weights_fine=w_f, warp_coarse=wf, warped_coarse=warped)
    return out


def as_tuple(out: Dict[str, torch.Tensor]):
    """Order the outputs like the reference pipelines return them."""
    if out['kind'] == 'smpl':
        return (out['rgb'], out['rgb_fine'], out['warp_out'], out['samples_out'], out['warped_out'], out['alpha_out'])
    return (out['rgb'], out['rgb_fine'], out['samples_out'], out['alpha_out'])


# --------------------------------------------------------------------------------------
# Weight sets used by fixtures and parity tests (SURVEY.md section 8c "fixtures")
# --------------------------------------------------------------------------------------
def build_nets(kind: str, seed: int, variant: str = 'default', *, n_layers=8, width=256, skips=(4,),
               L_pos=10, L_dir=4, L_pose=10, pose_encoded=True, net_cls=None, warp_cls=None, enc_cls=None):
    """Deterministically build (coarse, fine, warp|None, encoders) for pipeline ``kind``.

    ``net_cls`` / ``warp_cls`` / ``enc_cls`` let the caller substitute the REFERENCE classes:
    construction order and RNG consumption are identical, so the same seed gives the same
    weights for either class family (checked by tests/test_oracle_vs_reference.py).

    variant: 'default'  -- torch default init
             'dense'    -- default init, sigma head weight x20 and bias +1 (alpha spans 0..1)
             'sharp'    -- every parameter x2 (ill-conditioned; stresses the sampler)
    """
    net_cls = net_cls or RayNet
    warp_cls = warp_cls or WarpNet
    enc_cls = enc_cls or Encoder
    torch.manual_seed(seed)
    pos_enc, dir_enc, pose_enc = enc_cls(L_pos, False), enc_cls(L_dir, False), enc_cls(L_pose, False)
    P, D = 3 * pos_enc.output_dim, 3 * dir_enc.output_dim
    A = 0
    if kind == 'append':
        A = 2 * pose_enc.output_dim if pose_encoded else 2
    if kind == 'append_full':
        A = 69 * pose_enc.output_dim if pose_encoded else 69
    coarse = net_cls(n_layers, width, P, D, A, list(skips))
    fine = net_cls(n_layers, width, P, D, A, list(skips))
    warp = None
    if kind == 'smpl':
        warp = warp_cls(n_layers, width, P if pose_encoded else 3, 2 * pose_enc.output_dim if pose_encoded else 2)
    nets = [coarse, fine] + ([warp] if warp is not None else [])
    with torch.no_grad():
        if variant == 'dense':
            for net in (coarse, fine):
                net.sigma_out_layer.weight.mul_(20.)
                net.sigma_out_layer.bias.add_(1.)
        elif variant == 'sharp':
            for net in nets:
                for p in net.parameters():
                    p.mul_(2.)
        elif variant != 'default':
            raise ValueError(variant)
    for net in nets:
        net.eval()
    return coarse, fine, warp, pos_enc, dir_enc, pose_enc


def weight_checksum(nets: List[nn.Module]) -> float:
    """Order-sensitive fp64 checksum of all parameters (guards seed->weights reproducibility)."""
    s, k = 0.0, 1
    for net in nets:
        if net is None:
            continue
        for p in net.parameters():
            s += float(p.detach().double().sum()) * k + float(p.detach().double().abs().sum())
            k += 1
    return s
