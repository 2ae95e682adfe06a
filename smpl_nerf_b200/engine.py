"""Host side of the fused renderer: weight packing cache + the render call.

PyTorch is used for device memory, streams and nothing else; all arithmetic happens in
libnrf_b200.so (``nrf_render``: one persistent sm_100a kernel per ray batch).

Boundary being mirrored: the reference pipelines are constructed with live
``nn.Module`` nets whose parameters the optimizer mutates in place, so the packed (fp16 hi/lo,
UMMA-swizzled) copy is cached per net and keyed on every parameter's (data_ptr, _version); it is
re-packed on the device whenever a version changes.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import KIND, PipelineDesc, RayNetDesc, RenderIO, WarpNetDesc, check

POSE_COLS = (38, 41)     # the two SMPL arm angles the pipelines read (smpl_nerf_pipeline.py:28)


def _enc_cfg(enc):
    return int(enc.number_frequencies), 1 if enc.include_identity else 0


#: fold additional_linear_layer into the sigma head and directional_input at pack time (render_ray_net.py:51-57 has no
#: activation between them): one 256x256 layer less per sample, fp32-level reassociation.  On in both precision modes.
FOLD_LINEAR = True


def raynet_desc(net, pos_enc, dir_enc, per_sample_dirs: bool, ext_pose_bias: bool = False, fold: Optional[bool] = None) -> RayNetDesc:
    """Read the RenderRayNet hyper-parameters off the module (attribute names: render_ray_net.py:11-17)."""
    d = RayNetDesc()
    d.n_layers, d.width = int(net.n_layers), int(net.width)
    d.positions_dim = int(net.positions_dim)
    d.directions_dim = int(net.direcions_dim)       # (sic) reference spelling
    d.additional_input_dim = int(net.additional_input_dim)
    d.use_directional_input = 1 if net.use_directional_input else 0
    skips = [int(s) for s in net.skips if 0 <= int(s) < d.n_layers - 1]
    if len(skips) > _lib.NRF_MAX_SKIPS:
        raise ValueError(f'at most {_lib.NRF_MAX_SKIPS} skip connections are supported, got {skips}')
    d.n_skips = len(skips)
    for i, s in enumerate(skips):
        d.skips[i] = s
    d.pos_freqs, d.pos_identity = _enc_cfg(pos_enc)
    d.dir_freqs, d.dir_identity = _enc_cfg(dir_enc)
    d.per_sample_dirs = 1 if per_sample_dirs else 0
    d.ext_pose_bias = 1 if ext_pose_bias else 0
    d.fold_linear = 1 if (FOLD_LINEAR if fold is None else fold) else 0
    return d


def warpnet_desc(net, pos_enc, pose_dim: int, encoded: bool) -> WarpNetDesc:
    d = WarpNetDesc()
    d.width = int(net.linear1.out_features)
    d.positions_dim = int(net.linear1.in_features) - int(pose_dim)
    d.pose_dim = int(pose_dim)
    d.in_freqs, d.in_identity = _enc_cfg(pos_enc) if encoded else (0, 1)
    return d


class _Packed:
    __slots__ = ('storage', 'buf', 'key')


_cache: "weakref.WeakKeyDictionary[torch.nn.Module, _Packed]" = weakref.WeakKeyDictionary()


def invalidate(net=None) -> None:
    """Drop the packed copy of ``net`` (or of every net).  The cache key is every parameter's (data_ptr, _version):
    optimizer steps and load_state_dict bump the version, but in-place writes through ``p.data`` (``p.data.copy_()``, some
    EMA code) do NOT -- call this after such an update.  The packed buffer is re-filled in place on the stream of the next
    render call; renders of the same net from several streams at once must be ordered by the caller."""
    if net is None:
        _cache.clear()
    else:
        _cache.pop(net, None)


def _params(net, device) -> Sequence[torch.Tensor]:
    ps = [p.detach() for p in net.parameters()]
    for p in ps:
        if p.dtype != torch.float32:
            raise ValueError(f'net parameters must be float32, got {p.dtype}')
        if p.device != device:
            raise ValueError(f'net parameters live on {p.device} but the rays are on {device}')
    return [p if p.is_contiguous() else p.contiguous() for p in ps]


def _aligned_u8(nbytes: int, device) -> (torch.Tensor, torch.Tensor):
    storage = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-storage.data_ptr()) % 1024
    return storage, storage[off:off + nbytes]


def packed(net, desc, device, stream) -> torch.Tensor:
    """Device buffer holding ``net`` in the engine's layout; re-packed iff a parameter changed."""
    L = _lib.lib()
    # the hit path runs on every pipeline call: keep it to one walk over the parameters (~40 us for a RenderRayNet)
    key = (bytes(desc), device) + tuple([(p.data_ptr(), p._version) for p in net.parameters()])
    ent = _cache.get(net)
    if ent is not None and ent.key == key:
        return ent.buf
    ps = _params(net, device)
    is_warp = isinstance(desc, WarpNetDesc)
    nbytes = (L.nrf_warpnet_packed_bytes if is_warp else L.nrf_raynet_packed_bytes)(C.byref(desc))
    if nbytes == 0:
        check(-1, 'plan net')
    if ent is None or ent.buf.numel() != nbytes or ent.buf.device != device:
        ent = _Packed()
        ent.storage, ent.buf = _aligned_u8(nbytes, device)
    arr = (C.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
    fn = L.nrf_pack_warpnet if is_warp else L.nrf_pack_raynet
    check(fn(C.byref(desc), arr, len(ps), ent.buf.data_ptr(), stream), 'pack net')
    ent.key = key
    _cache[net] = ent
    return ent.buf


_u_cache: Dict = {}


def _u_fine(n_fine: int, device) -> torch.Tensor:
    k = (n_fine, str(device))
    if k not in _u_cache:
        # computed on the CPU so the bits match utils.py:206 run on the reference's CPU path
        _u_cache[k] = torch.linspace(0., 1., steps=n_fine, dtype=torch.float32).to(device)
    return _u_cache[k]


def _f32(t: torch.Tensor, name: str, device, shape=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise ValueError(f'{name} must be a tensor')
    if t.dtype != torch.float32:
        raise ValueError(f'{name} must be float32, got {t.dtype}')
    if t.device != device:
        raise ValueError(f'{name} is on {t.device}, expected {device}')
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f'{name} has shape {tuple(t.shape)}, expected {tuple(shape)}')
    return t.detach().contiguous()


def render(kind: str, model_coarse, model_fine, model_warp, args, pos_enc, dir_enc, pose_enc, data,
           *, taps: bool = False, z_all_in: Optional[torch.Tensor] = None, noise=None, n_sms: int = 0,
           precision: int = 0, trace_cap: int = 0, fold: Optional[bool] = None) -> Dict[str, torch.Tensor]:
    """One fused forward.  ``data`` is the reference's per-batch list
    [ray_samples, ray_translation, ray_direction, z_vals, (goal_pose,) rgb]; returns a dict of
    freshly allocated fp32 CUDA tensors (see NrfRenderIO in include/nrf_b200.h)."""
    L = _lib.lib()
    samples = data[0]
    device = samples.device
    if device.type != 'cuda':
        raise RuntimeError('smpl_nerf_b200 runs on CUDA devices only (no CPU fallback); got ' + str(device))
    if samples.dim() != 3 or samples.shape[-1] != 3:
        raise ValueError(f'ray_samples must be [B, n_coarse, 3], got {tuple(samples.shape)}')
    B, nc = int(samples.shape[0]), int(samples.shape[1])
    run_fine = 1 if args.run_fine else 0
    nf = int(args.number_fine_samples) if run_fine else 0
    n = nc + nf
    smpl = kind == 'smpl'
    samples = _f32(samples, 'ray_samples', device)
    origin = _f32(data[1], 'ray_translation', device, (B, 3))
    direction = _f32(data[2], 'ray_direction', device, (B, 3))
    z = _f32(data[3], 'z_vals', device, (B, nc))
    goal = None
    if kind != 'nerf':
        goal = _f32(data[4], 'goal_pose', device)
        if goal.dim() != 2 or goal.shape[0] != B or goal.shape[1] <= max(POSE_COLS):
            raise ValueError(f'goal_pose must be [B, >= {max(POSE_COLS) + 1}], got {tuple(goal.shape)}')
    full_pose = kind == 'append_full'     # AppendSmplParamsPipeline: all pose parameters, hoisted per ray by nrf_ray_bias

    # autograd is recording and a net is trainable: the differentiable layer-by-layer path (train.py) instead of the fused kernel
    from . import train as _train
    train_mode = _train.needs_grad(model_coarse, model_fine if run_fine else None, model_warp if smpl else None)
    # hidden widths other than 256 (config_parser.py:20,24,30 make netwidth a flag): the fused kernel is built for 256, the
    # layer-by-layer path (one tcgen05 GEMM per nn.Linear) takes 128 / 256 / 512 -- used for inference too in that case
    widths = [int(model_coarse.width)] + ([int(model_fine.width)] if run_fine else []) + \
        ([int(model_warp.linear1.out_features)] if smpl else [])
    if any(wd != 256 for wd in widths) or int(precision) == 2:
        # precision 2 = exact mode: bf16 hi/lo/ll planes (24 significant bits = exact fp32 operands, fp32's exponent range) and six
        # MMA passes per layer on the layer-by-layer path -- meets the literal 1e-4 raw-sigma bar on trained nets and has no
        # activation-range limit (the fused kernel's fp16 split carries 22 bits and saturates at 65504); inference only
        train_mode = True
    if train_mode and trace_cap > 0:
        raise ValueError('the trace tap exists in the fused inference kernel only')

    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        dc = raynet_desc(model_coarse, pos_enc, dir_enc, smpl, full_pose, fold)
        pc = None if train_mode else packed(model_coarse, dc, device, stream)
        df, pf = None, None
        if run_fine:
            df = raynet_desc(model_fine, pos_enc, dir_enc, smpl, full_pose, fold)
            pf = None if train_mode else packed(model_fine, df, device, stream)
        pipe = PipelineDesc()
        pipe.kind = KIND[kind]
        pipe.n_coarse, pipe.n_fine, pipe.run_fine = nc, nf, run_fine
        pipe.white_background = 1 if args.white_background else 0
        pipe.precision = int(precision)
        dw, pw = None, None
        ray_bias = []
        if full_pose and train_mode:
            pipe.pose_all = 1
            pipe.pose_freqs, pipe.pose_identity = _enc_cfg(pose_enc)
            pipe.pose_encoded = 1 if args.human_pose_encoding else 0
            pipe.pose_stride, pipe.pose_col0, pipe.pose_col1 = int(goal.shape[1]), 0, 0
        elif full_pose:
            # pose features exactly as models/append_smpl_params_pipeline.py:30-37 builds them, then one tcgen05 GEMM per net and hoisted layer
            encoded = bool(args.human_pose_encoding)
            A = int(goal.shape[1]) * (pose_enc.output_dim if encoded else 1)
            if A != dc.additional_input_dim:
                raise ValueError(f'pose feature count {A} does not match the net\'s additional_input_dim {dc.additional_input_dim}')
            feats = goal
            if encoded:
                feats = torch.empty(B, A, dtype=torch.float32, device=device)
                check(L.nrf_positional_encoding(goal.data_ptr(), B, int(goal.shape[1]), *_enc_cfg(pose_enc), feats.data_ptr(),
                                                stream), 'nrf_positional_encoding(goal_pose)')
            flag = torch.empty(1, dtype=torch.int32, device=device)     # 0 after nrf_ray_bias: one pose for the whole batch
            rb_out = []
            for net, desc in ((model_coarse, dc),) + (((model_fine, df),) if run_fine else ()):
                n_ext = L.nrf_raynet_ext_slots(C.byref(desc))
                if n_ext < 1:
                    check(-1, 'plan net')
                rb = torch.empty(B, n_ext, 256, dtype=torch.float32, device=device)
                ps = _params(net, device)
                arr = (C.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
                if B > 0:
                    wsb = L.nrf_ray_bias_workspace_bytes(C.byref(desc), B)
                    wsp = torch.empty(wsb + 256, dtype=torch.uint8, device=device)
                    woff = (-wsp.data_ptr()) % 256
                    check(L.nrf_ray_bias(C.byref(desc), arr, len(ps), feats.data_ptr(), B, rb.data_ptr(), flag.data_ptr(),
                                         wsp.data_ptr() + woff, wsb, stream), 'nrf_ray_bias')
                    ray_bias.append(wsp)
                rb_out.append(rb)
                ray_bias += [rb] + list(ps)          # keep (possibly re-laid-out) parameter tensors alive until the launch
            ray_bias += [feats, flag]
        elif kind != 'nerf':
            pipe.pose_freqs, pipe.pose_identity = _enc_cfg(pose_enc)
            pipe.pose_encoded = 1 if args.human_pose_encoding else 0
            pipe.pose_stride, pipe.pose_col0, pipe.pose_col1 = int(goal.shape[1]), POSE_COLS[0], POSE_COLS[1]
            pose_dim = 2 * (2 * pipe.pose_freqs + pipe.pose_identity) if pipe.pose_encoded else 2
            if smpl:
                dw = warpnet_desc(model_warp, pos_enc, pose_dim, bool(pipe.pose_encoded))
                pw = None if train_mode else packed(model_warp, dw, device, stream)

        io = RenderIO()
        out: Dict[str, torch.Tensor] = {}

        def new(name, *shape, dtype=torch.float32):
            t = torch.empty(shape, dtype=dtype, device=device)
            out[name] = t
            setattr(io, name, t.data_ptr())
            return t

        io.ray_samples, io.ray_origin, io.ray_dir, io.z_vals = (samples.data_ptr(), origin.data_ptr(),
                                                                direction.data_ptr(), z.data_ptr())
        if goal is not None:
            io.goal_pose = goal.data_ptr()
        keep = [samples, origin, direction, z, goal]
        if full_pose and not train_mode:
            io.ray_bias_nonuniform = flag.data_ptr()
            io.ray_bias_coarse = rb_out[0].data_ptr()
            if run_fine:
                io.ray_bias_fine = rb_out[1].data_ptr()
            keep += [t for t in ray_bias if isinstance(t, torch.Tensor)]
        if run_fine:
            u = _u_fine(nf, device)
            io.u_fine = u.data_ptr()
            keep.append(u)
        std = float(getattr(args, 'sigma_noise_std', 0.) or 0.)
        if noise is not None:
            n_c, n_f = noise
        elif std > 0.:
            # same draws, same order and shapes as utils.py:174 (coarse pass first, then fine)
            n_c = torch.normal(0, std, (B, nc), device=device)
            n_f = torch.normal(0, std, (B, n), device=device) if run_fine else None
        else:
            n_c = n_f = None
        if n_c is not None:
            n_c = _f32(n_c, 'noise_coarse', device, (B, nc)); io.noise_coarse = n_c.data_ptr(); keep.append(n_c)
        if n_f is not None and run_fine:
            n_f = _f32(n_f, 'noise_fine', device, (B, n)); io.noise_fine = n_f.data_ptr(); keep.append(n_f)
        if z_all_in is not None and run_fine:
            z_all_in = _f32(z_all_in, 'z_all_in', device, (B, n)); io.z_all_in = z_all_in.data_ptr(); keep.append(z_all_in)

        new('rgb', B, 3)
        if run_fine:
            new('rgb_fine', B, 3)
            new('samples_out', B, n, 3)
        else:
            out['rgb_fine'] = out['rgb']
            out['samples_out'] = data[0]
        new('alpha_out', B, n)
        if smpl:
            new('warp_out', B, n, 3)
            new('warped_out', B, n, 3)
        if taps:
            new('raw_coarse', B, nc, 4)
            new('weights_coarse', B, nc)
            if run_fine:
                new('raw_fine', B, n, 4)
                if not train_mode:
                    new('z_new', B, nf)          # (the layer-by-layer path does not expose the unmerged samples)
                new('z_all', B, n)
        if trace_cap > 0:       # developer tap: CTA 0's MMA / epilogue timeline
            tr = torch.zeros(2 + 3 * trace_cap, dtype=torch.int64, device=device)
            tr[0] = trace_cap
            out['trace'] = tr
            io.trace = tr.data_ptr()
        status = torch.zeros(1, dtype=torch.int32, device=device)
        out['status'] = status
        io.status = status.data_ptr()

        if train_mode:
            _train.render_train(kind, model_coarse, model_fine, model_warp, pipe, dc, df, dw, io, out, keep, B, device, n_sms)
        elif B > 0:
            rc = L.nrf_render(C.byref(pipe), C.byref(dc), pc.data_ptr(), C.byref(df) if df is not None else None,
                              pf.data_ptr() if pf is not None else None, C.byref(dw) if dw is not None else None,
                              pw.data_ptr() if pw is not None else None, C.byref(io), B, int(n_sms), stream)
            check(rc, 'nrf_render')
        for t in keep:      # the kernel is asynchronous: keep inputs alive on this stream
            if t is not None:
                t.record_stream(torch.cuda.current_stream(device))
    return out


def launches_per_render(kind: str = 'nerf', run_fine: bool = True, pose_encoded: bool = True) -> int:
    """Kernels of this library launched by one render() call with warm weight caches."""
    n = int(_lib.lib().nrf_render_launches())
    if kind == 'append_full':      # pose encoding + per net: plane split, row-uniformity probe, one tcgen05 GEMM per hoisted layer (2)
        n += (1 if pose_encoded else 0) + 4 * (2 if run_fine else 1)
    return n
