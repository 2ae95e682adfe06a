"""smpl_nerf_b200 -- B200-native volume-rendering engine behind the SMPL-NeRF pipeline API.

Drop-in replacements for the hot path of HannesStark/SMPL-NeRF (``models/*_pipeline.py`` forward,
``utils.py`` positional_encoding / raw2outputs / sample_pdf, ``torchsearchsorted``), implemented as
hand-written sm_100a CUDA behind a C ABI (include/nrf_b200.h -> csrc/libnrf_b200.so).

    from smpl_nerf_b200.models.nerf_pipeline import NerfPipeline
    from smpl_nerf_b200.models.smpl_nerf_pipeline import SmplNerfPipeline
    from smpl_nerf_b200.models.append_to_nerf_pipeline import AppendToNerfPipeline

There is no CPU or PyTorch fallback: without the built library or a CUDA device every call raises.
"""
from ._lib import build, lib  # noqa: F401

__all__ = ['build', 'lib']
