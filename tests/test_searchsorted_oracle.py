"""Pins the C restatement of torchsearchsorted's bisection: against numpy on the reference's own test
grid (torchsearchsorted/test/test_searchsorted.py:34-44, seeded here), against hand-made tie and
out-of-range vectors, and -- where oracle/_ref was built from /root/reference -- against the
reference's own compiled C++ extension.  Also checks that extension equals torch.searchsorted on the
hot-path shape (cdf[B,63], u = linspace(0,1,128)), which is why the oracle may use torch.searchsorted."""
import numpy as np
import pytest
import torch

from oracle import searchsorted_ref as S


def np_rowwise(a, v, side):
    rows = max(a.shape[0], v.shape[0])
    return np.stack([np.searchsorted(a[i if a.shape[0] > 1 else 0], v[i if v.shape[0] > 1 else 0], side=side)
                     for i in range(rows)])


@pytest.mark.parametrize('Ba,Bv', [(1, 1), (100, 100), (1, 100), (100, 1), (200, 200)])
@pytest.mark.parametrize('A', [1, 50, 500])
@pytest.mark.parametrize('V', [1, 12, 120])
@pytest.mark.parametrize('side', ['left', 'right'])
def test_c_oracle_matches_numpy(Ba, Bv, A, V, side):
    rng = np.random.RandomState(Ba * 7 + Bv * 3 + A + V)
    for _ in range(3):
        a = np.sort(rng.rand(Ba, A).astype(np.float32), -1)
        v = rng.rand(Bv, V).astype(np.float32)
        assert np.array_equal(S.c_oracle(a, v, side), np_rowwise(a, v, side))


def test_c_oracle_ties_and_range():
    a = np.array([[0., 1., 1., 1., 2., 3.]], np.float32)
    v = np.array([[-1., 0., 1., 1.5, 3., 4.]], np.float32)
    assert S.c_oracle(a, v, 'left').tolist() == [[0, 0, 1, 4, 5, 6]]
    assert S.c_oracle(a, v, 'right').tolist() == [[0, 1, 4, 4, 6, 6]]


@pytest.mark.skipif(not S.reference_available(), reason='oracle/_ref not built (reference tree absent)')
def test_compiled_reference_agrees():
    ref = S.reference_searchsorted()
    rng = np.random.RandomState(0)
    for (Ba, Bv, A, V) in [(1, 1, 1, 1), (100, 100, 50, 12), (1, 100, 500, 120), (100, 1, 50, 120), (64, 64, 63, 128)]:
        a = np.sort(rng.rand(Ba, A).astype(np.float32), -1)
        v = rng.rand(Bv, V).astype(np.float32)
        for side in ('left', 'right'):
            got = ref(torch.from_numpy(a), torch.from_numpy(v), side=side).numpy()
            assert np.array_equal(got, S.c_oracle(a, v, side))
            assert np.array_equal(got, np_rowwise(a, v, side))
    # hot-path shape: the oracle's torch.searchsorted(right=True) stand-in is index-identical
    torch.manual_seed(0)
    w = torch.rand(200, 62) ** 4 + 1e-5
    cdf = torch.cat([torch.zeros(200, 1), torch.cumsum(w / w.sum(-1, keepdim=True), -1)], -1)
    u = torch.linspace(0, 1, 128).expand(200, 128).contiguous()
    assert torch.equal(ref(cdf, u, side='right'), torch.searchsorted(cdf, u, right=True))


@pytest.mark.skipif(not S.reference_available(), reason='oracle/_ref not built (reference tree absent)')
def test_pipeline_with_compiled_reference_searchsorted(reference):
    """The whole reference pipeline run with ITS OWN compiled searchsorted equals the oracle."""
    from oracle import nerf_oracle as O
    from oracle import ref_import as R
    from smpl_nerf_b200 import scene
    ref2 = R.load(use_compiled_searchsorted=True)
    try:
        rays = scene.make_rays(8, 8, 64, seed=21)
        data = scene.data_list(rays, 'nerf')
        theirs = O.build_nets('nerf', 31, 'dense', net_cls=ref2.RenderRayNet, warp_cls=ref2.WarpFieldNet,
                              enc_cls=ref2.PositionalEncoder)
        mine = O.build_nets('nerf', 31, 'dense')
        args = O.make_args()
        with torch.no_grad():
            want = ref2.NerfPipeline(theirs[0], theirs[1], args, theirs[3], theirs[4])(data)
            got = O.as_tuple(O.nerf_forward(mine[0], mine[1], mine[3], mine[4], args, data))
        for a, b in zip(want, got):
            assert torch.equal(a, b)
    finally:
        R.load()   # rebind the torch.searchsorted stand-in
