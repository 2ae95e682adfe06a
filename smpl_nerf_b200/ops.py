"""Drop-in replacements for the reference's ``utils.py`` hot-path functions and for
``torchsearchsorted.searchsorted``, each a thin wrapper over one C-ABI call:

    utils.py:114-131  PositionalEncoder   -> PositionalEncoder (encode runs nrf_positional_encoding)
    utils.py:134-191  raw2outputs         -> raw2outputs
    utils.py:194-228  sample_pdf          -> sample_pdf
    utils.py:231-264  fine_sampling       -> fine_sampling
    torchsearchsorted/src/torchsearchsorted/searchsorted.py:20-53 -> searchsorted (same asserts)
    util/scores.py:88-173  ssim / gaussian_filter / img2mse / img2psnr -> ssim, gaussian_filter, img2mse, img2psnr

CUDA tensors only -- there is no CPU path in this package.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check
from .engine import _f32, _u_fine


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'smpl_nerf_b200.ops: {name} must be a CUDA tensor (no CPU fallback)')


class PositionalEncoder:
    """Same constructor/attributes as utils.py:114-125; ``encode`` is one CUDA kernel instead of
    2L pointwise launches and a cat."""

    def __init__(self, number_frequencies, include_identity):
        self.number_frequencies = int(number_frequencies)
        self.include_identity = include_identity
        self.output_dim = (1 if include_identity else 0) + 2 * self.number_frequencies

    def encode(self, coordinate: torch.Tensor) -> torch.Tensor:
        _need_cuda(coordinate, 'coordinate')
        if torch.is_grad_enabled() and coordinate.requires_grad:
            return _Encode.apply(coordinate, self.number_frequencies, 1 if self.include_identity else 0)
        return _encode_fwd(coordinate, self.number_frequencies, 1 if self.include_identity else 0)


def _encode_fwd(coordinate, freqs, identity):
    x = _f32(coordinate, 'coordinate', coordinate.device)
    c = int(x.shape[-1])
    n = x.numel() // c
    out = torch.empty(*x.shape[:-1], c * (identity + 2 * freqs), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().nrf_positional_encoding(x.data_ptr(), n, c, freqs, identity, out.data_ptr(), _stream(x.device)),
              'nrf_positional_encoding')
    return out


class _Encode(torch.autograd.Function):
    """Positional encoding with a native backward (nrf_positional_encoding_backward)."""

    @staticmethod
    def forward(ctx, x, freqs, identity):
        ctx.save_for_backward(x)
        ctx.cfg = (freqs, identity)
        return _encode_fwd(x, freqs, identity)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        freqs, identity = ctx.cfg
        xc = x.detach().contiguous().float()
        g = g.contiguous().float()
        c = int(xc.shape[-1])
        gx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            check(_lib.lib().nrf_positional_encoding_backward(xc.data_ptr(), g.data_ptr(), xc.numel() // c, c, freqs, identity,
                                                              gx.data_ptr(), _stream(xc.device)), 'nrf_positional_encoding_backward')
        return gx, None, None


def _raw2outputs_fwd(raw, z, dirs, noise, white):
    dev = raw.device
    B, n = int(raw.shape[0]), int(raw.shape[1])
    rgb = torch.empty(B, 3, dtype=torch.float32, device=dev)
    weights = torch.empty(B, n, dtype=torch.float32, device=dev)
    alpha = torch.empty(B, n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().nrf_raw2outputs(raw.data_ptr(), z.data_ptr(), dirs.data_ptr(),
                                         noise.data_ptr() if noise is not None else None, B, n, white, rgb.data_ptr(),
                                         weights.data_ptr(), alpha.data_ptr(), _stream(dev)), 'nrf_raw2outputs')
    return rgb, weights, alpha


class _Raw2Outputs(torch.autograd.Function):
    """raw2outputs with a native backward (nrf_raw2outputs_backward): gradients flow to ``raw`` only -- z_vals and the
    directions are data in every reference pipeline."""

    @staticmethod
    def forward(ctx, raw, z, dirs, noise, white):
        ctx.save_for_backward(raw, z, dirs, noise if noise is not None else raw.new_empty(0))
        ctx.white = white
        return _raw2outputs_fwd(raw, z, dirs, noise, white)

    @staticmethod
    def backward(ctx, g_rgb, g_w, g_a):
        raw, z, dirs, noise = ctx.saved_tensors
        dev = raw.device
        B, n = int(raw.shape[0]), int(raw.shape[1])
        g_raw = torch.empty_like(raw)
        g_rgb = torch.zeros(B, 3, dtype=torch.float32, device=dev) if g_rgb is None else g_rgb.contiguous().float()
        g_w = None if g_w is None else g_w.contiguous().float()
        g_a = None if g_a is None else g_a.contiguous().float()
        ptr = lambda t: t.data_ptr() if t is not None and t.numel() > 0 else None
        with torch.cuda.device(dev):
            check(_lib.lib().nrf_raw2outputs_backward(raw.data_ptr(), z.data_ptr(), dirs.data_ptr(), ptr(noise), B, n, ctx.white,
                                                      g_rgb.data_ptr(), ptr(g_w), ptr(g_a), g_raw.data_ptr(), _stream(dev)),
                  'nrf_raw2outputs_backward')
        return g_raw, None, None, None, None


def raw2outputs(raw: torch.Tensor, z_vals: torch.Tensor, samples_directions: torch.Tensor, args
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (rgb[B,3], weights[B,n], density(alpha)[B,n]); noise is drawn like utils.py:172-174.  Differentiable with
    respect to ``raw`` (native backward kernel) when ``raw.requires_grad``."""
    _need_cuda(raw, 'raw')
    dev = raw.device
    B, n = int(raw.shape[0]), int(raw.shape[1])
    needs_grad = torch.is_grad_enabled() and raw.requires_grad
    raw32 = raw if (needs_grad and raw.dtype == torch.float32 and raw.is_contiguous() and tuple(raw.shape) == (B, n, 4)) \
        else _f32(raw, 'raw', dev, (B, n, 4))
    z = _f32(z_vals, 'z_vals', dev, (B, n))
    dirs = _f32(samples_directions.expand(B, n, 3), 'samples_directions', dev, (B, n, 3))
    noise = None
    if float(getattr(args, 'sigma_noise_std', 0.) or 0.) > 0.:
        noise = torch.normal(0, float(args.sigma_noise_std), (B, n), device=dev)
    white = 1 if args.white_background else 0
    if needs_grad:
        if raw32 is not raw:
            raise ValueError('raw must be a contiguous float32 [B, n, 4] tensor to be differentiated')
        return _Raw2Outputs.apply(raw32, z, dirs, noise, white)
    return _raw2outputs_fwd(raw32, z, dirs, noise, white)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, args) -> torch.Tensor:
    _need_cuda(bins, 'bins')
    dev = bins.device
    B, m = int(bins.shape[0]), int(bins.shape[1])
    bins = _f32(bins, 'bins', dev)
    w = _f32(weights, 'weights', dev, (B, m - 1))
    nf = int(args.number_fine_samples)
    out = torch.empty(B, nf, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().nrf_sample_pdf(bins.data_ptr(), w.data_ptr(), _u_fine(nf, dev).data_ptr(), B, m, nf,
                                        out.data_ptr(), _stream(dev)), 'nrf_sample_pdf')
    return out


def fine_sampling(ray_translation: torch.Tensor, samples_directions: torch.Tensor, z_vals: torch.Tensor,
                  weights: torch.Tensor, args) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (z_vals[B, nc+nf] sorted, ray_samples_fine[B, nc+nf, 3])."""
    _need_cuda(z_vals, 'z_vals')
    dev = z_vals.device
    B, nc = int(z_vals.shape[0]), int(z_vals.shape[1])
    nf = int(args.number_fine_samples)
    o = _f32(ray_translation, 'ray_translation', dev, (B, 3))
    d = _f32(samples_directions, 'samples_directions', dev, (B, 3))
    z = _f32(z_vals, 'z_vals', dev)
    w = _f32(weights, 'weights', dev, (B, nc))
    z_all = torch.empty(B, nc + nf, dtype=torch.float32, device=dev)
    pts = torch.empty(B, nc + nf, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().nrf_fine_sampling(o.data_ptr(), d.data_ptr(), z.data_ptr(), w.data_ptr(),
                                           _u_fine(nf, dev).data_ptr(), B, nc, nf, z_all.data_ptr(), pts.data_ptr(),
                                           _stream(dev)), 'nrf_fine_sampling')
    return z_all, pts


def searchsorted(a: torch.Tensor, v: torch.Tensor, out: Optional[torch.Tensor] = None, side='left') -> torch.Tensor:
    """Same contract and assertions as torchsearchsorted.searchsorted (float32 CUDA inputs)."""
    assert len(a.shape) == 2, "input `a` must be 2-D."
    assert len(v.shape) == 2, "input `v` must be 2-D."
    assert (a.shape[0] == v.shape[0] or a.shape[0] == 1 or v.shape[0] == 1), \
        "`a` and `v` must have the same number of rows or one of them must have only one "
    assert a.device == v.device, '`a` and `v` must be on the same device'
    _need_cuda(a, 'a')
    result_shape = (max(a.shape[0], v.shape[0]), v.shape[1])
    if out is not None:
        assert out.device == a.device, "`out` must be on the same device as `a`"
        assert out.dtype == torch.long, "out.dtype must be torch.long"
        assert tuple(out.shape) == result_shape, "If the output tensor is provided, its shape must be correct."
        assert out.is_contiguous(), "`out` must be contiguous"
    else:
        out = torch.empty(result_shape, device=v.device, dtype=torch.long)
    # the reference silently returns garbage for non-contiguous inputs (its README says so); copy instead
    a32 = _f32(a, 'a', a.device)
    v32 = _f32(v, 'v', a.device)
    with torch.cuda.device(a.device):
        check(_lib.lib().nrf_searchsorted(a32.data_ptr(), a32.shape[0], a32.shape[1], v32.data_ptr(), v32.shape[0],
                                          v32.shape[1], out.data_ptr(), 1 if side == 'left' else 0,
                                          _stream(a.device)), 'nrf_searchsorted')
    return out


# --------------------------------------------------------------------------------------------------- metrics (util/scores.py)
def gaussian_filter(size: int, sigma: float) -> torch.Tensor:
    """2-D Gaussian window [1, size, size], built exactly like util/scores.py:68-86 (fp32 on the host)."""
    coords = torch.arange(size).to(dtype=torch.float32)
    coords -= (size - 1) / 2.
    g = coords ** 2
    g = (- (g.unsqueeze(0) + g.unsqueeze(1)) / (2 * sigma ** 2)).exp()
    g /= g.sum()
    return g.unsqueeze(0)


_window_cache = {}


def ssim(x: torch.Tensor, y: torch.Tensor, kernel_size: int = 11, kernel_sigma: float = 1.5, data_range=1., reduction: str = 'mean',
         full: bool = False, k1: float = 0.01, k2: float = 0.03):
    """Structural similarity with the signature and semantics of util/scores.py:88-131 for 4-D inputs ``(N, C, H, W)``
    (CUDA, fp32): per-channel valid-window SSIM (one kernel + a fixed-order reduction on the device), mean over channels, then
    ``reduction`` over the batch.  ``full=True`` also returns the contrast-structure term."""
    _need_cuda(x, 'x')
    if x.dim() != 4 or x.shape != y.shape:
        raise ValueError(f'ssim expects two (N, C, H, W) tensors of the same shape, got {tuple(x.shape)} and {tuple(y.shape)}')
    dev = x.device
    N, Cn, H, W = (int(v) for v in x.shape)
    if H < kernel_size or W < kernel_size:
        raise ValueError(f"Kernel size can't be greater than actual input size. Input size: {x.size()}. Kernel size: {kernel_size}")
    xx, yy = _f32(x, 'x', dev), _f32(y, 'y', dev)
    key = (kernel_size, float(kernel_sigma), str(dev))
    if key not in _window_cache:
        _window_cache[key] = gaussian_filter(kernel_size, kernel_sigma)[0].contiguous().to(dev)
    win = _window_cache[key]
    L = _lib.lib()
    planes = N * Cn
    partial = torch.empty(max(1, int(L.nrf_ssim_partial_floats(planes, H, W, kernel_size))), dtype=torch.float32, device=dev)
    s_out = torch.empty(N, Cn, dtype=torch.float32, device=dev)
    c_out = torch.empty(N, Cn, dtype=torch.float32, device=dev)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    with torch.cuda.device(dev):
        check(L.nrf_ssim(xx.data_ptr(), yy.data_ptr(), planes, H, W, win.data_ptr(), kernel_size, c1, c2, partial.data_ptr(),
                         s_out.data_ptr(), c_out.data_ptr(), _stream(dev)), 'nrf_ssim')
    ssim_val, cs = s_out.mean(1), c_out.mean(1)
    if reduction != 'none':
        op = {'mean': torch.mean, 'sum': torch.sum}[reduction]
        ssim_val, cs = op(ssim_val, dim=0), op(cs, dim=0)
    return (ssim_val, cs) if full else ssim_val


def img2mse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return torch.mean((x - y) ** 2)           # util/scores.py:11-28


def img2psnr(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return -10. * torch.log10(torch.mean((x - y) ** 2))       # util/scores.py:30-48 (log / log(10))


def to_uint8_bgr(images: torch.Tensor, to_bgr: bool = True) -> torch.Tensor:
    """inference.py:260-262 on the device: ``clip(img, 0, 1) * 255`` -> uint8 (truncating cast), channels flipped to BGR."""
    _need_cuda(images, 'images')
    img = _f32(images, 'images', images.device)
    if img.shape[-1] != 3:
        raise ValueError('images must be [..., 3]')
    out = torch.empty(img.shape, dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        check(_lib.lib().nrf_quantize_image(img.data_ptr(), img.numel() // 3, out.data_ptr(), 1 if to_bgr else 0, _stream(img.device)),
              'nrf_quantize_image')
    return out
