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


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, 'smpl_nerf_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(dirpath, fn)
