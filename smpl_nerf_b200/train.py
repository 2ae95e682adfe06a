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


# One training workspace per device is kept between steps (a solver loop asks for the same ~1.2 MB per ray every step; handing it back
# to PyTorch's caching allocator each time let any >1 MB allocation made between two steps -- e.g. the next batch's host->device copy --
# split the cached block, and the following step then paid a multi-GB cudaMalloc).  A forward whose graph is still alive owns the cached
# buffer; a second forward issued before that backward gets a private allocation.  The buffer is reused in stream order: issue the
# steps of one device from one stream (what the reference's solvers do).
_ws_cache: Dict[int, list] = {}      # device index -> [uint8 tensor, in use]


def _ws_acquire(device, nbytes: int):
    ent = _ws_cache.get(device.index)
    if ent is not None and not ent[1] and ent[0].numel() >= nbytes:
        ent[1] = True
        return ent[0], True
    t = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if ent is None or not ent[1]:
        _ws_cache[device.index] = [t, True]
        return t, True
    return t, False


def _ws_release(device, t) -> None:
    ent = _ws_cache.get(device.index)
    if ent is not None and ent[0] is t:
        ent[1] = False


def release_workspaces() -> None:
    """Drop the cached training workspaces (they are otherwise kept until the process ends)."""
    for k in [k for k, ent in _ws_cache.items() if not ent[1]]:
        del _ws_cache[k]


class _Call:
    """Everything one forward/backward pair shares (descs, io struct, tensors kept alive)."""
    __slots__ = ('pipe', 'dc', 'df', 'dw', 'io', 'B', 'n_c', 'n_f', 'n_w', 'keep', 'rgb_bufs', 'workspace', 'ws_cached', 'device', 'smpl', 'run_fine', 'n_sms')

    def drop_workspace(self) -> None:
        if self.workspace is not None and self.ws_cached:
            _ws_release(self.device, self.workspace)
        self.workspace = None

    def __del__(self):          # a graph that is never back-propagated (loss only evaluated) frees the cached buffer too
        try:
            self.drop_workspace()
        except Exception:
            pass


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
            call.workspace, call.ws_cached = _ws_acquire(call.device, ws_bytes + 256)
            off = (-call.workspace.data_ptr()) % 256
            rc = L.nrf_train_forward(C.byref(call.pipe), C.byref(call.dc), _ptr_table(pc), len(pc),
                                     C.byref(call.df) if call.df is not None else None, _ptr_table(pf) if pf else None, len(pf),
                                     C.byref(call.dw) if call.dw is not None else None, _ptr_table(pw) if pw else None, len(pw),
                                     C.byref(call.io), call.B, call.workspace.data_ptr() + off, ws_bytes, call.n_sms, stream)
            check(rc, 'nrf_train_forward')
        ctx.call = call
        ctx.save_for_backward(*params)
        # the colour buffers become the graph-carrying outputs: the call must not keep them (tensor -> grad_fn -> ctx -> call would be a
        # reference cycle through C++ that Python's collector cannot see; every step would leak its outputs and its workspace)
        res, call.rgb_bufs = call.rgb_bufs, None
        return tuple(res)

    @staticmethod
    def backward(ctx, *gs):
        call: _Call = ctx.call
        L = _lib.lib()
        if call.workspace is None:
            raise RuntimeError('smpl_nerf_b200: backward through this pipeline call ran already (its saved activations are released after the '
                               'first backward; retain_graph is not supported)')
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
        call.drop_workspace()        # one backward per forward (like retain_graph=False)
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
    call.rgb_bufs = [out['rgb']] + ([out['rgb_fine']] if run_fine else [])
    call.keep = {'out': {k: v for k, v in out.items() if k not in ('rgb', 'rgb_fine')}, 'inputs': keep,
                 'rgb_storage': [t.detach() for t in call.rgb_bufs]}      # graph-less aliases: the io struct points into them
    call.workspace, call.ws_cached = None, False
    flat = [p for ps in params for p in ps]
    res = RenderTrain.apply(call, *flat)
    out['rgb'] = res[0]
    out['rgb_fine'] = res[1] if run_fine else res[0]
