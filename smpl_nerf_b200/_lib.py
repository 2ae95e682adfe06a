"""ctypes binding of libnrf_b200.so (the C ABI declared in include/nrf_b200.h).

The library is built in-tree with nvcc for sm_100a (``build()``; also run by
``__graft_entry__.build()``).  There is NO fallback: if the shared object is missing or a call
fails, a RuntimeError is raised -- the product path never computes on the CPU or through the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.environ.get('NRF_LIB_PATH') or os.path.join(CSRC, 'libnrf_b200.so')   # NRF_LIB_PATH: A/B builds (developer)
SOURCES = ['nrf_pack.cu', 'nrf_fused.cu', 'nrf_ops.cu', 'nrf_diag.cu', 'nrf_gemm.cu', 'nrf_train.cu']
HEADERS = ['nrf_plan.h', 'nrf_ptx.cuh', 'nrf_stages.cuh', 'nrf_gemm.cuh', 'nrf_train_kernels.cuh', os.path.join('..', '..', 'include', 'nrf_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']

NRF_MAX_SKIPS = 4
ABI_VERSION = 4          # NRF_ABI_VERSION of include/nrf_b200.h this binding was written against
KIND = {'nerf': 0, 'smpl': 1, 'append': 2, 'append_full': 2}   # append_full = append + externally hoisted 69-parameter pose


class RayNetDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('width', C.c_int32), ('positions_dim', C.c_int32),
                ('directions_dim', C.c_int32), ('additional_input_dim', C.c_int32),
                ('use_directional_input', C.c_int32), ('n_skips', C.c_int32),
                ('skips', C.c_int32 * NRF_MAX_SKIPS), ('pos_freqs', C.c_int32), ('pos_identity', C.c_int32),
                ('dir_freqs', C.c_int32), ('dir_identity', C.c_int32), ('per_sample_dirs', C.c_int32),
                ('ext_pose_bias', C.c_int32), ('fold_linear', C.c_int32)]


class WarpNetDesc(C.Structure):
    _fields_ = [('width', C.c_int32), ('positions_dim', C.c_int32), ('pose_dim', C.c_int32),
                ('in_freqs', C.c_int32), ('in_identity', C.c_int32)]


class PipelineDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('n_coarse', C.c_int32), ('n_fine', C.c_int32), ('run_fine', C.c_int32),
                ('white_background', C.c_int32), ('pose_freqs', C.c_int32), ('pose_identity', C.c_int32),
                ('pose_encoded', C.c_int32), ('pose_stride', C.c_int32), ('pose_col0', C.c_int32),
                ('pose_col1', C.c_int32), ('precision', C.c_int32), ('pose_all', C.c_int32)]


_IO_IN = ['ray_samples', 'ray_origin', 'ray_dir', 'z_vals', 'goal_pose', 'u_fine', 'noise_coarse', 'noise_fine',
          'z_all_in', 'ray_bias_coarse', 'ray_bias_fine', 'ray_bias_nonuniform']
_IO_OUT = ['rgb', 'rgb_fine', 'samples_out', 'alpha_out', 'warp_out', 'warped_out', 'raw_coarse', 'raw_fine',
           'weights_coarse', 'z_new', 'z_all', 'status', 'trace']


class RenderIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _IO_IN + _IO_OUT]


EXPORTS = ['nrf_last_error', 'nrf_abi_version', 'nrf_device_supported', 'nrf_raynet_packed_bytes',
           'nrf_warpnet_packed_bytes', 'nrf_pack_raynet', 'nrf_pack_warpnet', 'nrf_render', 'nrf_render_launches',
           'nrf_raynet_ext_slots', 'nrf_ray_bias', 'nrf_ray_bias_workspace_bytes', 'nrf_generate_rays', 'nrf_generate_rays_range',
           'nrf_ssim', 'nrf_ssim_partial_floats', 'nrf_quantize_image',
           'nrf_positional_encoding', 'nrf_positional_encoding_backward', 'nrf_raw2outputs', 'nrf_raw2outputs_backward', 'nrf_sample_pdf', 'nrf_fine_sampling', 'nrf_searchsorted',
           'nrf_selftest_umma', 'nrf_selftest_umma2', 'nrf_bench_umma', 'nrf_bench_umma2',
           'nrf_train_workspace_bytes', 'nrf_train_forward', 'nrf_train_backward', 'nrf_split_planes', 'nrf_gemm_planes', 'nrf_gemm_dw', 'nrf_train_launch_count']


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(nvcc):
        raise RuntimeError('nvcc not found: cannot build libnrf_b200.so (and there is no CPU fallback)')
    extra = os.environ.get('NRF_NVCC_EXTRA', '').split()     # e.g. -DNRF_SOME_EXPERIMENT=1 (developer A/B builds)
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (never builds implicitly on import; call build() first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(smpl_nerf_b200 has no CPU / PyTorch fallback)')
    L = C.CDLL(LIB_PATH)
    L.nrf_last_error.restype = C.c_char_p
    L.nrf_abi_version.restype = C.c_int
    L.nrf_device_supported.argtypes = [C.c_int]
    L.nrf_raynet_packed_bytes.restype = C.c_size_t
    L.nrf_raynet_packed_bytes.argtypes = [C.POINTER(RayNetDesc)]
    L.nrf_warpnet_packed_bytes.restype = C.c_size_t
    L.nrf_warpnet_packed_bytes.argtypes = [C.POINTER(WarpNetDesc)]
    L.nrf_pack_raynet.argtypes = [C.POINTER(RayNetDesc), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]
    L.nrf_pack_warpnet.argtypes = [C.POINTER(WarpNetDesc), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]
    L.nrf_render.argtypes = [C.POINTER(PipelineDesc), C.POINTER(RayNetDesc), C.c_void_p, C.POINTER(RayNetDesc),
                             C.c_void_p, C.POINTER(WarpNetDesc), C.c_void_p, C.POINTER(RenderIO), C.c_int64, C.c_int,
                             C.c_void_p]
    L.nrf_raynet_ext_slots.argtypes = [C.POINTER(RayNetDesc)]
    L.nrf_ray_bias.argtypes = [C.POINTER(RayNetDesc), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]
    L.nrf_ray_bias_workspace_bytes.restype = C.c_size_t
    L.nrf_ray_bias_workspace_bytes.argtypes = [C.POINTER(RayNetDesc), C.c_int64]
    L.nrf_generate_rays.argtypes = [C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_generate_rays_range.argtypes = [C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_ssim_partial_floats.restype = C.c_int64
    L.nrf_ssim_partial_floats.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32]
    L.nrf_ssim.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_quantize_image.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
    L.nrf_positional_encoding.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.nrf_raw2outputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_positional_encoding_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                   C.c_void_p]
    L.nrf_raw2outputs_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_sample_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p]
    L.nrf_fine_sampling.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_searchsorted.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_int32, C.c_void_p]
    L.nrf_selftest_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_bench_umma2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    L.nrf_selftest_umma2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_bench_umma.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    PP = C.POINTER(C.c_void_p)
    L.nrf_train_workspace_bytes.restype = C.c_size_t
    L.nrf_train_workspace_bytes.argtypes = [C.POINTER(PipelineDesc), C.POINTER(RayNetDesc), C.POINTER(RayNetDesc), C.POINTER(WarpNetDesc), C.c_int64]
    _net_args = [C.POINTER(PipelineDesc), C.POINTER(RayNetDesc), PP, C.c_int, C.POINTER(RayNetDesc), PP, C.c_int, C.POINTER(WarpNetDesc), PP, C.c_int,
                 C.POINTER(RenderIO), C.c_int64, C.c_void_p, C.c_size_t]
    L.nrf_train_forward.argtypes = _net_args + [C.c_int, C.c_void_p]
    L.nrf_train_backward.argtypes = _net_args + [C.c_void_p, C.c_void_p, PP, PP, PP, C.c_int, C.c_void_p]
    L.nrf_train_launch_count.restype = C.c_longlong
    L.nrf_train_launch_count.argtypes = [C.c_int]
    L.nrf_split_planes.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.nrf_gemm_planes.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nrf_gemm_dw.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                              C.c_void_p, C.c_void_p]
    if L.nrf_abi_version() != ABI_VERSION:
        raise RuntimeError('libnrf_b200.so ABI version mismatch; rebuild')
    _lib = L
    return L


def check(rc: int, what: str = 'nrf call') -> None:
    """Map a non-zero C return code to an exception (ValueError for argument errors)."""
    if rc == 0:
        return
    msg = lib().nrf_last_error().decode(errors='replace')
    if rc == -1:
        raise ValueError(f'{what}: {msg}')
    raise RuntimeError(f'{what} failed ({rc}): {msg}')


def exported_symbols() -> List[str]:
    return list(EXPORTS)
