"""tcgen05.mma issue-rate probe (developer tool): cycles per K=64 step for N=128/256, with and
without a concurrent TMA weight stream, on 1 SM and on all 148."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smpl_nerf_b200 import _lib
L = _lib.lib()
w = torch.zeros(8 << 20, dtype=torch.uint8, device='cuda')
iters = 4000
for n_ctas in (148,):
    for mode in (1, 3, 5, 7):
        cyc = torch.zeros(n_ctas, dtype=torch.int64, device='cuda')
        _lib.check(L.nrf_bench_umma(mode, iters, w.data_ptr(), w.numel(), cyc.data_ptr(), n_ctas, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        n = 256 if mode & 1 else 128
        passes = 2 if mode & 4 else 1
        floor = 4 * 128 * n / 256 * passes
        c = cyc.double() / iters
        print(f'ctas={n_ctas:3d} N={n} tma={bool(mode & 2)!s:5} a_passes={passes}: {c.mean():8.1f} cyc/step (max {c.max():8.1f}; floor {floor:.0f}) '
              f'-> {100 * floor / c.mean():5.1f}% of the tensor floor', flush=True)

print('--- CTA pairs (cta_group::2, M=256, N=256, 16 KB half-stage per CTA; one step = one [256 x 64] stage)')
for n_pairs in (74,):
    for mode, slots in ((0, 3), (4, 3), (1, 3), (5, 3), (3, 3), (7, 3)):
        cyc = torch.zeros(n_pairs, dtype=torch.int64, device='cuda')
        _lib.check(L.nrf_bench_umma2(mode, iters, slots, w.data_ptr(), w.numel(), cyc.data_ptr(), n_pairs, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        passes = 2 if mode & 4 else 1
        floor = 4 * (64 if mode & 1 else 128) * passes
        c = cyc.double() / iters
        print(f'N={128 if mode & 1 else 256} pairs={n_pairs:3d} {"random-data" if mode & 32 else "ones"} {"16-warps-parked" if mode & 64 else ""} tma={bool(mode & 2)!s:5} slots={slots} a_passes={passes}: {c.mean():8.1f} cyc/stage '
              f'(max {c.max():8.1f}; floor {floor}) -> {100 * floor / c.mean():5.1f}% of the tensor floor; {16384 / c.mean():5.1f} B/clk/SM streamed', flush=True)
