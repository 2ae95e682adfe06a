"""Stand-alone ``forward`` of the two MLPs on pre-encoded features (models/render_ray_net.py:42-61, models/warp_field_net.py:17-21).

The pipelines never call this (they run the nets inside the fused kernel or the training engine); it exists so that the drop-in
``RenderRayNet`` / ``WarpFieldNet`` classes are callable like the reference's for code that evaluates a net directly.  Every
``nn.Linear`` is one launch of the library's tcgen05 GEMM (``nrf_gemm_planes``: fp16 hi/lo operand planes, three passes, fp32
accumulate, bias + ReLU fused) -- PyTorch only allocates, concatenates and pads.  Inference only; CUDA only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _split(x: torch.Tensor, rows_pad: int, cols_pad: int):
    """fp32 [rows, cols] -> zero-padded fp16 (hi, lo) planes [rows_pad, cols_pad]."""
    L = _lib.lib()
    rows, cols = x.shape
    hi = torch.zeros(rows_pad, cols_pad, dtype=torch.float16, device=x.device)
    lo = torch.zeros(rows_pad, cols_pad, dtype=torch.float16, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    check(L.nrf_split_planes(x.data_ptr(), rows, cols, x.stride(0), hi.data_ptr(), lo.data_ptr(), None, cols_pad, cols_pad, C.c_void_p(stream)), 'nrf_split_planes')
    return hi, lo


def linear(x: torch.Tensor, layer: nn.Linear, relu: bool) -> torch.Tensor:
    """[S, in] fp32 -> [S, out] fp32 = (relu)(x W^T + b) on the tensor cores."""
    L = _lib.lib()
    S, K = x.shape
    N = layer.out_features
    if K != layer.in_features:
        raise ValueError(f'linear: input has {K} features, the layer takes {layer.in_features}')
    Kp, Np = _pad64(K), _pad64(N)
    with torch.cuda.device(x.device):
        xh, xl = _split(x.contiguous(), S, Kp)
        wh, wl = _split(layer.weight.detach().contiguous(), Np, Kp)
        bias = torch.zeros(Np, dtype=torch.float32, device=x.device)
        if layer.bias is not None:
            bias[:N] = layer.bias.detach()
        out = torch.empty(S, Np, dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        check(L.nrf_gemm_planes(0, xh.data_ptr(), xl.data_ptr(), None, S, Kp, wh.data_ptr(), wl.data_ptr(), None, Np, 3, bias.data_ptr(),
                                1 if relu else 0, out.data_ptr(), None, None, None, C.c_void_p(stream)), 'nrf_gemm_planes')
    return out[:, :N]


def _prepare(net: nn.Module, x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError('smpl_nerf_b200: the nets run on CUDA tensors only (there is no CPU fallback)')
    if torch.is_grad_enabled() and net.training and any(p.requires_grad for p in net.parameters()):
        raise NotImplementedError('the stand-alone net forward is inference-only; train through a smpl_nerf_b200.models.*Pipeline '
                                  '(call .eval() or torch.no_grad() to evaluate the net alone)')
    lead = x.shape[:-1]
    return x.reshape(-1, x.shape[-1]).float(), lead


def render_ray_net_forward(net, x: torch.Tensor) -> torch.Tensor:
    """models/render_ray_net.py:42-61: [..., positions + additional (+ ...) + directions] -> [..., 4] = (rgb_raw, sigma_raw)."""
    x2, lead = _prepare(net, x)
    first_in = net.positions_dim + net.additional_input_dim
    pp, dirs = x2[:, :first_in], x2[:, x2.shape[1] - net.direcions_dim:]
    o = linear(pp, net.positions_pose_input, True)
    for i, layer in enumerate(net.positional_net):
        o = linear(torch.cat([o, pp], -1) if i in net.skips else o, layer, True)
    o = linear(o, net.additional_linear_layer, False)
    sigma = linear(o, net.sigma_out_layer, False)
    o = linear(torch.cat([o, dirs], -1) if net.use_directional_input else o, net.directional_input, False)
    for layer in net.directional_net:
        o = linear(o, layer, True)
    rgb = linear(o, net.rgb_out_layer, False)
    return torch.cat([rgb, sigma], -1).reshape(*lead, 4)


def warp_field_net_forward(net, x: torch.Tensor) -> torch.Tensor:
    """models/warp_field_net.py:17-21: [..., positions + pose] -> [..., 3]."""
    x2, lead = _prepare(net, x)
    return linear(linear(x2, net.linear1, True), net.linear2, False).reshape(*lead, 3)
