"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the public header
declares, the planner accepts/rejects shapes with the documented error behaviour, and the product
path refuses to run without a CUDA device (no CPU fallback).  No GPU compute is launched here."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import nerf_oracle as O
from smpl_nerf_b200 import _lib, engine
from smpl_nerf_b200.models import AppendToNerfPipeline, NerfPipeline, SmplNerfPipeline
from smpl_nerf_b200.models.singe_sample_pipeline import SmplPipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    _lib.build()
    return _lib.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, 'include', 'nrf_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(nrf_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/nrf_b200.h but not exported'
    assert set(_lib.exported_symbols()) == names
    assert L.nrf_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header(L):
    # int32 counts in the C structs (include/nrf_b200.h)
    assert C.sizeof(_lib.RayNetDesc) == 4 * (7 + 4 + 7)
    assert C.sizeof(_lib.WarpNetDesc) == 4 * 5
    assert C.sizeof(_lib.PipelineDesc) == 4 * 13
    assert C.sizeof(_lib.RenderIO) == 8 * 25


def test_planner_sizes_and_rejections(L):
    c, f, w, pe, de, he = O.build_nets('nerf', 0)
    d = engine.raynet_desc(c, pe, de, False)
    n = L.nrf_raynet_packed_bytes(C.byref(d))
    # 2 x fp16 (hi, lo) copy of every MMA weight, padded K: about 2.4 MB
    assert 2_300_000 < n < 2_700_000 and n % 1024 == 0
    d.width = 128
    assert L.nrf_raynet_packed_bytes(C.byref(d)) == 0
    assert b'width' in L.nrf_last_error()
    d = engine.raynet_desc(c, pe, de, False)
    d.positions_dim = 63
    assert L.nrf_raynet_packed_bytes(C.byref(d)) == 0
    assert b'positions_dim' in L.nrf_last_error()
    # NULL parameter table -> NRF_E_INVALID, message set, nothing launched
    d = engine.raynet_desc(c, pe, de, False)
    assert L.nrf_pack_raynet(C.byref(d), None, 0, None, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, 'x')
    with pytest.raises(RuntimeError):
        _lib.check(-2, 'x')


def test_descs_read_module_attributes():
    c, f, w, pe, de, he = O.build_nets('append', 0)
    d = engine.raynet_desc(c, pe, de, False)
    assert (d.n_layers, d.width, d.positions_dim, d.directions_dim, d.additional_input_dim) == (8, 256, 60, 24, 40)
    assert d.n_skips == 1 and d.skips[0] == 4 and (d.pos_freqs, d.dir_freqs) == (10, 4)
    c, f, w, pe, de, he = O.build_nets('smpl', 0)
    dw = engine.warpnet_desc(w, pe, 40, True)
    assert (dw.width, dw.positions_dim, dw.pose_dim, dw.in_freqs, dw.in_identity) == (256, 60, 40, 10, 0)


def test_pipeline_constructors_mirror_reference():
    c, f, w, pe, de, he = O.build_nets('smpl', 0)
    args = O.make_args()
    p1 = NerfPipeline(c, f, args, pe, de)
    p2 = AppendToNerfPipeline(c, f, args, pe, de, he)
    p3 = SmplNerfPipeline(c, f, w, args, pe, de, he)
    for p in (p1, p2, p3):
        assert isinstance(p, SmplPipeline) and isinstance(p, torch.nn.Module)
        assert p.model_coarse is c and p.model_fine is f and p.args is args
        assert p.position_encoder is pe and p.direction_encoder is de
    assert p3.model_warp_field is w and p3.human_pose_encoder is he
    # nets are registered sub-modules: they show up in parameters() like in the reference
    assert len(list(p3.parameters())) == len(list(c.parameters())) * 2 + 4


def test_no_cpu_fallback():
    from smpl_nerf_b200 import scene
    c, f, w, pe, de, he = O.build_nets('nerf', 0)
    data = scene.data_list(scene.make_rays(4, 4, 64), 'nerf')
    pipe = NerfPipeline(c, f, O.make_args(), pe, de)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pipe(data)


def test_ops_and_ray_generation_refuse_cpu_tensors():
    import numpy as np
    from types import SimpleNamespace
    from smpl_nerf_b200 import ops, rays, scene
    args = SimpleNamespace(sigma_noise_std=0., white_background=1, number_fine_samples=8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.PositionalEncoder(4, False).encode(torch.zeros(3, 3))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.raw2outputs(torch.zeros(2, 8, 4), torch.zeros(2, 8), torch.zeros(2, 8, 3), args)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.sample_pdf(torch.zeros(2, 7), torch.zeros(2, 6), args)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.searchsorted(torch.zeros(1, 4), torch.zeros(1, 2))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        rays.generate_view(4, 4, scene.sphere_pose(0., 0.), device='cpu')
    with pytest.raises(ValueError):
        rays.generate_view(4, 4, np.eye(3), device='cuda:0')
    from smpl_nerf_b200.models import RenderRayNet, WarpFieldNet          # the stand-alone net forward: CUDA tensors only
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        RenderRayNet(4, 128, 60, 24, 0, [2]).eval()(torch.zeros(5, 84))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        WarpFieldNet(8, 256, 60, 40).eval()(torch.zeros(5, 100))


def test_render_argument_errors_are_reported_without_a_gpu(L):
    """nrf_render validates its descriptors before it touches the device: NRF_E_INVALID (-1) + a message."""
    c, f, w, pe, de, he = O.build_nets('smpl', 0)
    dc = engine.raynet_desc(c, pe, de, True)
    io = _lib.RenderIO()
    pipe = _lib.PipelineDesc()
    pipe.kind, pipe.n_coarse, pipe.n_fine, pipe.run_fine = 1, 64, 128, 1
    blob = C.c_void_p(1024)            # aligned dummy address: never dereferenced on these paths
    assert L.nrf_render(None, C.byref(dc), blob, None, None, None, None, C.byref(io), 8, 0, None) == -1
    assert L.nrf_render(C.byref(pipe), C.byref(dc), blob, None, None, None, None, C.byref(io), -1, 0, None) == -1
    assert b'n_rays' in L.nrf_last_error()
    assert L.nrf_render(C.byref(pipe), C.byref(dc), blob, None, None, None, None, C.byref(io), 0, 0, None) == 0     # empty batch: nothing to do
    assert L.nrf_render(C.byref(pipe), C.byref(dc), blob, None, None, None, None, C.byref(io), 8, 0, None) == -1
    assert b'fine net' in L.nrf_last_error()
    pipe.run_fine = 0
    assert L.nrf_render(C.byref(pipe), C.byref(dc), blob, None, None, None, None, C.byref(io), 8, 0, None) == -1
    assert b'warp net' in L.nrf_last_error()
    pipe.kind = 7
    assert L.nrf_render(C.byref(pipe), C.byref(dc), blob, None, None, None, None, C.byref(io), 8, 0, None) == -1
    assert b'kind' in L.nrf_last_error()
    assert L.nrf_ray_bias(C.byref(dc), None, 0, None, 8, None, None, None, 0, None) == -1          # not an ext_pose_bias net
    assert b'ext_pose_bias' in L.nrf_last_error()
    assert L.nrf_generate_rays(0, 4, 1.0, None, None, None, None, 8, None, None, None, None, None) == -1


def test_training_entry_points_validate_without_a_gpu(L):
    """nrf_train_* / nrf_ssim / nrf_generate_rays_range / nrf_ray_bias check shapes, pointers and workspace sizes before any launch."""
    c, f, w, pe, de, he = O.build_nets('smpl', 0)
    dc, dw = engine.raynet_desc(c, pe, de, True), engine.warpnet_desc(w, pe, 40, True)
    pipe = _lib.PipelineDesc()
    pipe.kind, pipe.n_coarse, pipe.n_fine, pipe.run_fine, pipe.white_background = 1, 64, 128, 1, 1
    pipe.pose_freqs, pipe.pose_identity, pipe.pose_encoded, pipe.pose_stride, pipe.pose_col0, pipe.pose_col1 = 10, 0, 1, 69, 38, 41
    small = L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(dc), C.byref(dc), C.byref(dw), 64)
    big = L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(dc), C.byref(dc), C.byref(dw), 1024)
    # ~17 KB of saved activations / gradient planes per sample (256 samples per ray) + fixed buffers
    assert 0 < small < big and 2.5e9 < big < 6e9
    pipe.precision = 1          # one-pass mode keeps no lo planes
    assert L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(dc), C.byref(dc), C.byref(dw), 1024) < 0.8 * big
    pipe.precision = 0
    dbad = engine.raynet_desc(c, pe, de, True); dbad.width = 192
    assert L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(dbad), C.byref(dc), C.byref(dw), 64) == 0
    assert b'width' in L.nrf_last_error()
    d128 = engine.raynet_desc(c, pe, de, True); d128.width = 128      # 128 / 512 are planned by the layer-by-layer path
    assert L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(d128), C.byref(d128), C.byref(dw), 64) > 0
    pipe.pose_encoded = 0
    assert L.nrf_train_workspace_bytes(C.byref(pipe), C.byref(dc), C.byref(dc), C.byref(dw), 64) == 0
    assert b'human_pose_encoding' in L.nrf_last_error()
    pipe.pose_encoded = 1
    io = _lib.RenderIO()
    PP = C.POINTER(C.c_void_p)
    tab = (C.c_void_p * 26)(*([1024] * 26))
    tw = (C.c_void_p * 4)(*([1024] * 4))
    args = (C.byref(pipe), C.byref(dc), tab, 26, C.byref(dc), tab, 26, C.byref(dw), tw, 4)
    assert L.nrf_train_forward(*args, None, 8, C.c_void_p(4096), big, 0, None) == -1 and b'io is NULL' in L.nrf_last_error()
    assert L.nrf_train_forward(*args, C.byref(io), 8, None, big, 0, None) == -1 and b'workspace' in L.nrf_last_error()
    assert L.nrf_train_forward(*args, C.byref(io), 8, C.c_void_p(4096), 1024, 0, None) == -1 and b'too small' in L.nrf_last_error()
    assert L.nrf_train_forward(C.byref(pipe), C.byref(dc), tab, 25, C.byref(dc), tab, 26, C.byref(dw), tw, 4, C.byref(io), 8, C.c_void_p(4096), big, 0,
                               None) == -1 and b'parameter tensors' in L.nrf_last_error()
    assert L.nrf_train_forward(*args, C.byref(io), 8, C.c_void_p(4096), big, 0, None) == -1 and b'ray inputs' in L.nrf_last_error()
    # image metrics / ray windows / per-ray bias
    assert L.nrf_ssim(C.c_void_p(16), C.c_void_p(16), 3, 32, 32, C.c_void_p(16), 10, 1e-4, 9e-4, C.c_void_p(16), C.c_void_p(16), None, None) == -1
    assert L.nrf_ssim(C.c_void_p(16), C.c_void_p(16), 3, 8, 32, C.c_void_p(16), 11, 1e-4, 9e-4, C.c_void_p(16), C.c_void_p(16), None, None) == -1
    assert b"Kernel size can't be greater" in L.nrf_last_error()            # the reference's message (util/scores.py:147-149)
    assert L.nrf_ssim_partial_floats(3, 64, 64, 11) == 3 * 2 * 16
    cam = (C.c_double * 16)(*([0.0] * 16))
    ptr = C.c_void_p(16)
    assert L.nrf_generate_rays_range(8, 8, 1.0, cam, ptr, ptr, ptr, 4, 60, 8, ptr, ptr, ptr, ptr, None) == -1 and b'window' in L.nrf_last_error()
    ca, fa, *_ = O.build_nets('append_full', 0)
    dfull = engine.raynet_desc(ca, pe, de, False, True)
    assert L.nrf_raynet_ext_slots(C.byref(dfull)) == 2
    # fp16 hi/lo planes of 4096 feature rows and of two 256-row weight blocks, K = 1380 padded to 1408
    assert L.nrf_ray_bias_workspace_bytes(C.byref(dfull), 4096) >= (4096 + 2 * 256) * 1408 * 4
    assert L.nrf_train_launch_count(1) >= 0 and L.nrf_train_launch_count(0) == 0


def test_grad_mode_routing_rule():
    """train.needs_grad: differentiable path iff autograd records AND a net is in training mode AND a parameter requires grad."""
    from smpl_nerf_b200 import train
    c, f, w, *_ = O.build_nets('smpl', 0)          # build_nets leaves the nets in eval mode
    assert not train.needs_grad(c, f, w)
    c.train()
    assert train.needs_grad(c, f, w) and train.needs_grad(c, None, None)
    with torch.no_grad():
        assert not train.needs_grad(c, f, w)
    for p in c.parameters():
        p.requires_grad_(False)
    assert train.needs_grad(c, f, w)               # fine / warp parameters still require grad
    for m in (f, w):
        for p in m.parameters():
            p.requires_grad_(False)
    assert not train.needs_grad(c, f, w)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, 'smpl_nerf_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(dirpath, fn)
