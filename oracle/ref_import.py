"""Import the ACTUAL reference hot path from /root/reference -- TEST INFRASTRUCTURE ONLY.

Recipe from SURVEY.md section 8(c): the reference's ``utils.py`` imports five packages that are
not installed here (imageio, matplotlib, trimesh, mpl_toolkits, torchsearchsorted); none is
used on the hot path, so empty module stubs are enough, and ``torch.searchsorted(right=...)``
stands in for ``torchsearchsorted.searchsorted`` (index-identical; when oracle/_ref holds the
reference's own compiled C++ searchsorted, ``use_compiled_searchsorted=True`` binds that
instead).  The reference tree is never modified and nothing is copied out of it.

Only usable where /root/reference exists: the build container.  The GPU box never sees it;
tests that need it skip there and rely on the committed fixtures in tests/golden/.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get('NRF_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'nerf_pipeline.py'))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_cached = None


def load(use_compiled_searchsorted: bool = False):
    """Returns a namespace with the reference classes/functions of the hot path."""
    global _cached
    if _cached is not None and not use_compiled_searchsorted:
        return _cached
    if not available():
        raise RuntimeError(f'reference tree not found under {REFERENCE_ROOT}')

    def _searchsorted(a, v, out=None, side='left'):
        res = torch.searchsorted(a, v, right=(side != 'left'))
        if out is not None:
            out.copy_(res)
            return out
        return res

    if use_compiled_searchsorted:
        from . import searchsorted_ref
        _searchsorted = searchsorted_ref.reference_searchsorted()

    _stub('imageio')
    _stub('matplotlib')
    _stub('matplotlib.pyplot')
    _stub('mpl_toolkits')
    _stub('mpl_toolkits.axes_grid1', make_axes_locatable=None)
    tm = _stub('trimesh')
    tm.base = types.SimpleNamespace(Trimesh=object)
    _stub('trimesh.ray')
    _stub('trimesh.ray.ray_triangle', RayMeshIntersector=object)
    tss = _stub('torchsearchsorted')
    tss.searchsorted = _searchsorted

    # the reference uses top-level module names ("utils", "models"); make sure OUR packages of
    # the same names are not already imported under those names
    for name in ('utils', 'models'):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        # `models` is a namespace package in the reference (no __init__.py): it has __path__, no __file__
        where = getattr(mod, '__file__', None) or ''.join(list(getattr(mod, '__path__', []))[:1])
        if not str(where).startswith(REFERENCE_ROOT):
            raise RuntimeError(f'module name clash: {name} already imported from {where}')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import utils as ref_utils                                            # noqa: E402
    ref_utils.searchsorted = _searchsorted
    from models.render_ray_net import RenderRayNet                       # noqa: E402
    from models.warp_field_net import WarpFieldNet                       # noqa: E402
    from models.nerf_pipeline import NerfPipeline                        # noqa: E402
    from models.smpl_nerf_pipeline import SmplNerfPipeline               # noqa: E402
    from models.append_to_nerf_pipeline import AppendToNerfPipeline      # noqa: E402
    from models.append_smpl_params_pipeline import AppendSmplParamsPipeline  # noqa: E402

    ns = types.SimpleNamespace(
        utils=ref_utils, PositionalEncoder=ref_utils.PositionalEncoder,
        raw2outputs=ref_utils.raw2outputs, sample_pdf=ref_utils.sample_pdf,
        fine_sampling=ref_utils.fine_sampling, RenderRayNet=RenderRayNet, WarpFieldNet=WarpFieldNet,
        NerfPipeline=NerfPipeline, SmplNerfPipeline=SmplNerfPipeline,
        AppendToNerfPipeline=AppendToNerfPipeline, AppendSmplParamsPipeline=AppendSmplParamsPipeline)
    if not use_compiled_searchsorted:
        _cached = ns
    return ns


def build_pipeline(kind: str, coarse, fine, warp, pos_enc, dir_enc, pose_enc, args):
    ref = load()
    if kind == 'nerf':
        return ref.NerfPipeline(coarse, fine, args, pos_enc, dir_enc)
    if kind == 'append':
        return ref.AppendToNerfPipeline(coarse, fine, args, pos_enc, dir_enc, pose_enc)
    if kind == 'append_full':
        return ref.AppendSmplParamsPipeline(coarse, fine, args, pos_enc, dir_enc, pose_enc)
    if kind == 'smpl':
        return ref.SmplNerfPipeline(coarse, fine, warp, args, pos_enc, dir_enc, pose_enc)
    raise ValueError(kind)
