"""Error table of the CUDA engine against the reference's outputs stored in tests/golden/*.pt (developer tool; the
oracle package is used as the checker only).  For every fixture and both precision modes:
  sigma_raw / rgb_raw (coarse net, and fine net teacher-forced with the reference's depths), alpha, rgb, rgb_fine
next to the reference's OWN fp32-vs-fp64 deviation on the same inputs (the numerical noise floor)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smpl_nerf_b200 import engine
from tests import helpers as H

DEV = 'cuda:0'


def stats(a, b):
    e = (a.detach().cpu().double() - b.double()).abs().flatten()
    return f'{float(e.max()):.1e}/{float(torch.quantile(e, 0.999)) if e.numel() > 1 else float(e.max()):.1e}'


print('max / 99.9th-percentile absolute error; tolerances: sigma, alpha 1e-4, rgb 1e-3')
print(f'{"fixture":22s} {"mode":7s} {"sigma_c":>15s} {"rgbraw_c":>15s} {"rgb":>15s} {"sigma_f(tf)":>15s} {"alpha(tf)":>15s} '
      f'{"rgb_fine(tf)":>15s} {"rgb_fine(free)":>15s} | {"ref fp32-fp64 rgb_fine":>22s}')
for name in H.fixtures():
    fx = H.load_fixture(name)
    kind = fx['kind']
    nets = H.nets_for(fx)
    args = H.args_for(fx)
    gnets, gdata = H.to_cuda(nets, fx['data'])
    c, f, w, pe, de, he = gnets
    inter, ref = fx['intermediates'], fx['reference_outputs']
    floor = stats(ref[1], fx['reference_outputs_fp64'][1])
    for mode, prec in (('parity', 0), ('fast', 1)):
        got = engine.render(kind, c, f, w, args, pe, de, he, gdata, taps=True, precision=prec)
        cols = [stats(got['raw_coarse'][..., 3], inter['raw_coarse'][..., 3]), stats(got['raw_coarse'][..., :3], inter['raw_coarse'][..., :3]),
                stats(got['rgb'], ref[0])]
        if fx['run_fine']:
            tf = engine.render(kind, c, f, w, args, pe, de, he, gdata, taps=True, precision=prec, z_all_in=inter['z_all'].to(DEV))
            sig = inter['raw_fine'][..., 3]
            mask = H.alpha_mask_well_conditioned(sig)
            ea = (tf['alpha_out'].cpu().double() - ref[-1].double()).abs()[mask]
            cols += [stats(tf['raw_fine'][..., 3], sig), f'{float(ea.max()):.1e}/{float(torch.quantile(ea, 0.999)):.1e}',
                     stats(tf['rgb_fine'], ref[1]), stats(got['rgb_fine'], ref[1])]
        else:
            cols += ['-', stats(got['alpha_out'], ref[-1]), '-', '-']
        print(f'{name:22s} {mode:7s} ' + ' '.join(f'{x:>15s}' for x in cols) + f' | {floor:>22s}')
ck, nets, args = H.load_trained()
gnets, gdata = H.to_cuda(nets, ck['data'])
c, f, _, pe, de, he = gnets
for mode, prec in (('parity', 0), ('fast', 1)):
    got = engine.render('nerf', c, f, None, args, pe, de, he, gdata, precision=prec)
    print(f'trained_nerf_d4 (held-out 32x32) {mode:7s} rgb {stats(got["rgb"], ck["reference_rgb"])}  rgb_fine {stats(got["rgb_fine"], ck["reference_rgb_fine"])}  '
          f'PSNR vs GT {H.psnr(got["rgb_fine"], ck["data"][-1]):.4f} dB (reference {ck["reference_psnr"]:.4f} dB)')
