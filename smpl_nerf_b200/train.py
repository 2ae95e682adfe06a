"""Differentiable pipeline call: what makes ``loss.backward()`` of the reference's training loops work.

    solver/nerf_solver.py:81-87        out = self.pipeline(data); loss = MSE(out[0]) + MSE(out[1]); loss.backward(); optim.step()
    solver/smpl_nerf_solver.py:74-83   the same with the warp-field net's parameters in the optimizer

When autograd is recording and a net parameter requires grad, ``engine.render`` routes here instead of the fused
inference kernel: ``nrf_train_forward`` evaluates the pipeline layer by layer (one tcgen05 GEMM per ``nn.Linear``,
everything the backward needs kept in a workspace) and ``nrf_train_backward`` produces d(loss)/d(parameter) for every
``nn.Linear`` of the coarse, fine and warp nets -- all in libnrf_b200.so; PyTorch only owns the buffers and the autograd
node.  ``rgb`` and ``rgb_fine`` are differentiable; the sample points, alpha ("densities") and warp outputs are returned
without a graph (the reference's solvers do not differentiate them unless the optional GMM density loss is switched on,
which this path does not support; the hierarchical sampler is detached in the reference too, utils.py:260)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import KIND, PipelineDesc, RenderIO, check


def needs_grad(*nets) -> bool:
    """True when this call must be differentiable: autograd is recording, a parameter requires grad AND a net is in
    training mode.  The solvers put the nets in ``.train()`` for training batches and in ``.eval()`` for validation
    (solver/nerf_solver.py:73-74, 94-95, 115-116); ``inference.py:247-254`` and the solvers' early validation call the
    pipeline with eval-mode nets WITHOUT ``torch.no_grad()`` and never call backward -- those stay on the fused kernel."""
    if not torch.is_grad_enabled():
        return False
    nets = [n for n in nets if n is not None]
    if not any(getattr(n, 'training', True) for n in nets):
        return False
    return any(p.requires_grad for n in nets for p in n.parameters())


def _ptr_table(tensors: List[torch.Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _Call:
    """Everything one forward/backward pair shares (descs, io struct, tensors kept alive)."""
    __slots__ = ('pipe', 'dc', 'df', 'dw', 'io', 'B', 'n_c', 'n_f', 'n_w', 'keep', 'workspace', 'device', 'smpl', 'run_fine', 'n_sms')


class RenderTrain(torch.autograd.Function):

    @staticmethod
    def forward(ctx, call: _Call, *params: torch.Tensor):
        L = _lib.lib()
        ps = [p.detach() for p in params]
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != call.device:
                raise ValueError('trainable net parameters must be contiguous float32 tensors on the rays\' device')
        pc, pf, pw = ps[:call.n_c], ps[call.n_c:call.n_c + call.n_f], ps[call.n_c + call.n_f:]
        with torch.cuda.device(call.device):
            stream = torch.cuda.current_stream(call.device).cuda_stream
            ws_bytes = L.nrf_train_workspace_bytes(C.byref(call.pipe), C.byref(call.dc), C.byref(call.df) if call.df is not None else None,
                                                   C.byref(call.dw) if call.dw is not None else None, call.B)
            if ws_bytes == 0:
                check(-1, 'nrf_train_workspace_bytes')
            call.workspace = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=call.device)
            off = (-call.workspace.data_ptr()) % 256
            rc = L.nrf_train_forward(C.byref(call.pipe), C.byref(call.dc), _ptr_table(pc), len(pc),
                                     C.byref(call.df) if call.df is not None else None, _ptr_table(pf) if pf else None, len(pf),
                                     C.byref(call.dw) if call.dw is not None else None, _ptr_table(pw) if pw else None, len(pw),
                                     C.byref(call.io), call.B, call.workspace.data_ptr() + off, ws_bytes, call.n_sms, stream)
            check(rc, 'nrf_train_forward')
        ctx.call = call
        ctx.save_for_backward(*params)
        out = call.keep['out']
        res = [out['rgb']] + ([out['rgb_fine']] if call.run_fine else [])
        return tuple(res)

    @staticmethod
    def backward(ctx, *gs):
        call: _Call = ctx.call
        L = _lib.lib()
        params = ctx.saved_tensors
        ps = [p.detach() for p in params]
        pc, pf, pw = ps[:call.n_c], ps[call.n_c:call.n_c + call.n_f], ps[call.n_c + call.n_f:]
        grads = [torch.zeros_like(p) for p in ps]
        gc, gf, gw = grads[:call.n_c], grads[call.n_c:call.n_c + call.n_f], grads[call.n_c + call.n_f:]
        zero3 = lambda: torch.zeros(call.B, 3, dtype=torch.float32, device=call.device)
        g_rgb = gs[0].contiguous().float() if gs[0] is not None else zero3()
        g_fine = None
        if call.run_fine:
            g_fine = gs[1].contiguous().float() if gs[1] is not None else zero3()
        with torch.cuda.device(call.device):
            stream = torch.cuda.current_stream(call.device).cuda_stream
            off = (-call.workspace.data_ptr()) % 256
            rc = L.nrf_train_backward(C.byref(call.pipe), C.byref(call.dc), _ptr_table(pc), len(pc),
                                      C.byref(call.df) if call.df is not None else None, _ptr_table(pf) if pf else None, len(pf),
                                      C.byref(call.dw) if call.dw is not None else None, _ptr_table(pw) if pw else None, len(pw),
                                      C.byref(call.io), call.B, call.workspace.data_ptr() + off, call.workspace.numel() - 256,
                                      g_rgb.data_ptr(), g_fine.data_ptr() if g_fine is not None else None,
                                      _ptr_table(gc), _ptr_table(gf) if gf else None, _ptr_table(gw) if gw else None, call.n_sms, stream)
            check(rc, 'nrf_train_backward')
        call.workspace = None        # one backward per forward (like retain_graph=False)
        return (None,) + tuple(grads)


def render_train(kind: str, model_coarse, model_fine, model_warp, pipe: PipelineDesc, dc, df, dw, io: RenderIO, out: Dict[str, torch.Tensor],
                 keep: list, B: int, device, n_sms: int = 0) -> None:
    """Run the differentiable path; ``out['rgb']`` / ``out['rgb_fine']`` are replaced by graph-carrying tensors."""
    run_fine = bool(pipe.run_fine)
    nets = [model_coarse] + ([model_fine] if run_fine else []) + ([model_warp] if kind == 'smpl' else [])
    if any(net is None for net in nets):
        raise ValueError('a net required by this pipeline is None')
    params = [list(net.parameters()) for net in nets]
    call = _Call()
    call.pipe, call.dc, call.df, call.dw, call.io, call.B, call.device = pipe, dc, df if run_fine else None, dw, io, B, device
    call.n_c = len(params[0])
    call.n_f = len(params[1]) if run_fine else 0
    call.n_w = len(params[-1]) if kind == 'smpl' else 0
    call.smpl, call.run_fine, call.n_sms = kind == 'smpl', run_fine, int(n_sms)
    call.keep = {'out': out, 'inputs': keep}
    call.workspace = None
    flat = [p for ps in params for p in ps]
    res = RenderTrain.apply(call, *flat)
    out['rgb'] = res[0]
    out['rgb_fine'] = res[1] if run_fine else res[0]
