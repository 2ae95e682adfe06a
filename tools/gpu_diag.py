"""First-contact GPU diagnostic: tcgen05 self-test, then stage-wise parity of each pipeline kind
against the CPU oracle on a small batch.  Run under `timeout` on the GPU box.  (Uses oracle/ as the
checker only -- this is a developer tool, not product code.)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_oracle as O          # noqa: E402
from smpl_nerf_b200 import _lib, engine, scene   # noqa: E402


def p(*a):
    print(*a, flush=True)


def selftest():
    L = _lib.lib()
    torch.manual_seed(0)
    a = torch.randn(128, 64, device='cuda')
    b = torch.randn(256, 64, device='cuda')
    d = torch.zeros(128, 256, device='cuda')
    _lib.check(L.nrf_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.half().float() @ b.half().float().t()
    err = (d - ref).abs().max().item()
    p('selftest umma max err', err, 'ref absmax', ref.abs().max().item())
    return err < 1e-2


def run_kind(kind, variant='dense', B=70, run_fine=1, n_coarse=64, n_fine=128, n_layers=8, skips=(4,), pose_encoded=True,
             precision=0, seed=1):
    nets = O.build_nets(kind, seed, variant, n_layers=n_layers, skips=skips, pose_encoded=pose_encoded)
    c, f, w, pe, de, he = nets
    rays = scene.make_rays(16, 16, n_coarse, seed=seed)
    args = O.make_args(run_fine=run_fine, number_fine_samples=n_fine, human_pose_encoding=1 if pose_encoded else 0)
    data = scene.data_list(rays, kind, slice(0, B))
    with torch.no_grad():
        if kind == 'nerf':
            want = O.nerf_forward(c, f, pe, de, args, data)
        elif kind == 'append':
            want = O.append_to_nerf_forward(c, f, pe, de, he, args, data)
        else:
            want = O.smpl_nerf_forward(c, f, w, pe, de, he, args, data)
    dev = torch.device('cuda:0')
    gc, gf = c.to(dev), f.to(dev)
    gw = w.to(dev) if w is not None else None
    gdata = [t.to(dev) for t in data]
    t0 = time.time()
    got = engine.render(kind, gc, gf, gw, args, pe, de, he, gdata, taps=True, precision=precision)
    torch.cuda.synchronize()
    p(f'[{kind}/{variant} B={B} fine={run_fine}] render ok in {time.time() - t0:.3f}s status={int(got["status"].item())}')

    def cmp(name, a, b):
        a = a.cpu()
        e = (a - b).abs()
        p(f'   {name:16s} max|d|={e.max().item():.3e} mean|d|={e.mean().item():.3e}  ref absmax={b.abs().max().item():.3e}')
        return e.max().item()

    cmp('raw_coarse', got['raw_coarse'], want['raw_coarse'])
    cmp('weights_coarse', got['weights_coarse'], want['weights_coarse'])
    cmp('rgb', got['rgb'], want['rgb'])
    if run_fine:
        cmp('z_new', got['z_new'], want['z_new'])
        cmp('z_all', got['z_all'], want['z_all'])
        cmp('samples_out', got['samples_out'], want['samples_out'])
        cmp('raw_fine', got['raw_fine'], want['raw_fine'])
        cmp('rgb_fine', got['rgb_fine'], want['rgb_fine'])
    cmp('alpha_out', got['alpha_out'], want['alpha_out'])
    if kind == 'smpl':
        cmp('warp_out', got['warp_out'], want['warp_out'])
        cmp('warped_out', got['warped_out'], want['warped_out'])


if __name__ == '__main__':
    p(torch.cuda.get_device_name(0))
    ok = selftest()
    if not ok:
        p('SELFTEST FAILED -- stopping')
        sys.exit(1)
    which = sys.argv[1:] or ['nerf0', 'nerf', 'append', 'smpl']
    if 'nerf0' in which:
        run_kind('nerf', B=8, run_fine=0)
    if 'nerf' in which:
        run_kind('nerf')
    if 'append' in which:
        run_kind('append')
    if 'smpl' in which:
        run_kind('smpl')
    if 'cfg1' in which:
        run_kind('nerf', B=100, run_fine=0, n_coarse=32, n_layers=4, skips=())
    p('DIAG DONE')
