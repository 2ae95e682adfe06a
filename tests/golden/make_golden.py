"""Mint golden fixtures from the ACTUAL reference (imported from /root/reference, build container only).

    python tests/golden/make_golden.py

For every (pipeline kind, weight variant) case this stores, in tests/golden/<case>.pt:
  * the seeded inputs (the reference's ``data`` list) and how the nets were built (seed, variant,
    hyper-parameters, an fp64 checksum of all weights -- default-init weights are regenerated from the
    seed by ``oracle.nerf_oracle.build_nets`` instead of being stored, 2.4 MB per net),
  * the reference pipeline's own outputs (fp32, run through the reference classes),
  * the same forward in fp64 (reference classes cast to double) = the numerical noise floor,
  * stage intermediates (raw, weights, z_new, z_all) from the oracle restatement, which this script
    asserts to be bit-identical to the reference on the final outputs first.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import ref_import as R       # noqa: E402
from smpl_nerf_b200 import scene         # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, kind, variant, seed, B, kwargs
    ('nerf_dense', 'nerf', 'dense', 101, 16, {}),
    ('nerf_sharp', 'nerf', 'sharp', 102, 16, {}),
    ('append_dense', 'append', 'dense', 103, 16, {}),
    ('append_rawpose', 'append', 'dense', 104, 12, dict(pose_encoded=False)),
    ('smpl_dense', 'smpl', 'dense', 105, 16, {}),
    ('smpl_default', 'smpl', 'default', 106, 12, {}),
    ('nerf_cfg1', 'nerf', 'dense', 107, 20, dict(n_layers=4, skips=(), n_coarse=32, run_fine=0)),
    ('smpl_coarse_rawpose', 'smpl', 'dense', 108, 12, dict(pose_encoded=False, run_fine=0)),
    # AppendSmplParamsPipeline: all 69 pose parameters, per-ray random poses
    ('append_full_dense', 'append_full', 'dense', 109, 12, {}),
    ('append_full_rawpose', 'append_full', 'dense', 110, 12, dict(pose_encoded=False)),
]


def run_oracle(kind, nets, args, data):
    c, f, w, pe, de, he = nets
    if kind == 'nerf':
        return O.nerf_forward(c, f, pe, de, args, data)
    if kind == 'append':
        return O.append_to_nerf_forward(c, f, pe, de, he, args, data)
    if kind == 'append_full':
        return O.append_smpl_params_forward(c, f, pe, de, he, args, data)
    return O.smpl_nerf_forward(c, f, w, pe, de, he, args, data)


def main():
    ref = R.load()
    only = set(sys.argv[1:])
    for name, kind, variant, seed, B, kw in CASES:
        if only and name not in only:
            continue
        kw = dict(kw)
        n_coarse = kw.pop('n_coarse', 64)
        run_fine = kw.pop('run_fine', 1)
        pose_encoded = kw.get('pose_encoded', True)
        build = dict(kw)
        rays = scene.make_rays(24, 24, n_coarse, seed=seed, phi=15.0, theta=40.0 + seed, arm_angle_deg=(seed * 7) % 60)
        # spread the B rays over the image so that they hit different depths/directions
        sel = torch.linspace(0, 24 * 24 - 1, B).long()
        data = scene.data_list(rays, kind, sel)
        if kind == 'append_full':      # every ray gets its own full 69-parameter pose (|angle| < 1 rad)
            g = torch.Generator().manual_seed(seed)
            data[4] = (torch.rand(data[4].shape, generator=g) * 2 - 1).float()
        args = O.make_args(run_fine=run_fine, human_pose_encoding=1 if pose_encoded else 0)
        theirs = O.build_nets(kind, seed, variant, net_cls=ref.RenderRayNet, warp_cls=ref.WarpFieldNet,
                              enc_cls=ref.PositionalEncoder, **build)
        mine = O.build_nets(kind, seed, variant, **build)
        with torch.no_grad():
            want = R.build_pipeline(kind, *theirs, args)(data)
            inter = run_oracle(kind, mine, args, data)
            got = O.as_tuple(inter)
            assert all(torch.equal(a, b) for a, b in zip(want, got)), name
            # fp64 noise floor: the reference classes themselves, cast to double
            nets64 = [n.double() if n is not None else None for n in theirs[:3]]
            data64 = [t.double() for t in data]
            want64 = R.build_pipeline(kind, nets64[0], nets64[1], nets64[2], theirs[3], theirs[4], theirs[5], args)(data64)
        fixture = dict(
            name=name, kind=kind, variant=variant, seed=seed, build=build, n_coarse=n_coarse, run_fine=run_fine,
            pose_encoded=pose_encoded, torch_version=torch.__version__,
            weight_checksum=O.weight_checksum(list(mine[:3])),
            data=[t.clone() for t in data],
            reference_outputs=[t.clone() for t in want],
            reference_outputs_fp64=[t.clone() for t in want64],
            intermediates={k: v.clone() for k, v in inter.items()
                           if k in ('raw_coarse', 'weights_coarse', 'alpha_coarse', 'z_new', 'z_all', 'raw_fine')},
        )
        path = os.path.join(OUT, name + '.pt')
        torch.save(fixture, path)
        d32 = [float((a.double() - b).abs().max()) for a, b in zip(want, want64)]
        print(f'{name:22s} {os.path.getsize(path) / 1024:7.1f} KB   fp32-vs-fp64 max|d| per output: ' +
              ' '.join(f'{x:.1e}' for x in d32))


if __name__ == '__main__':
    main()
