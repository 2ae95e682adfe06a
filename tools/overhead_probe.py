"""Host-side cost of one pipeline call: wall time per call at tiny / reference-default batch sizes (developer tool)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smpl_nerf_b200.models import SmplNerfPipeline
w = bench.WORKLOADS['cfg2']
coarse, fine, warp, pe, de, he = bench.build_models(w)
dev = torch.device('cuda:0')
coarse, fine, warp = coarse.to(dev), fine.to(dev), warp.to(dev)
pipe = SmplNerfPipeline(coarse, fine, warp, bench.make_args(w), pe, de, he)
full = [t.to(dev) for t in bench.make_views(w, 0, 1)[0]]
for B in (2, 296, 800, 4096, 16384):
    data = [t[:B].contiguous() for t in full]
    with torch.no_grad():
        for _ in range(5):
            pipe(data)
        torch.cuda.synchronize()
        n = 200 if B <= 4096 else 30
        t0 = time.perf_counter()
        for _ in range(n):
            out = pipe(data)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print(f'B={B:6d}: {1e6 * t_issue / n:8.1f} us/call host issue, {1e6 * t_all / n:8.1f} us/call incl. GPU  -> {B * n / t_all:12.0f} rays/s')

# --- where does the 16,384-ray call spend its time when the caller holds the previous outputs?
data = full
import torch.cuda as tc
for mode in ('discard', 'hold', 'hold+stats'):
    with torch.no_grad():
        for _ in range(3):
            pipe(data)
        tc.synchronize()
        tc.reset_peak_memory_stats()
        n0 = tc.memory_stats()['num_device_alloc'] if 'num_device_alloc' in tc.memory_stats() else -1
        e0, e1 = tc.Event(enable_timing=True), tc.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        keep = None
        for _ in range(20):
            o = pipe(data)
            if mode != 'discard':
                keep = o
            del o
        e1.record()
        t_issue = time.perf_counter() - t0
        tc.synchronize()
        st = tc.memory_stats()
    print(f'{mode:11s}: host issue {1e3 * t_issue / 20:6.2f} ms/call, GPU {e0.elapsed_time(e1) / 20:6.2f} ms/call, '
          f"cudaMalloc calls so far {st.get('num_device_alloc', -1)}, frees {st.get('num_device_free', -1)}, peak {tc.max_memory_allocated() / 2**20:.0f} MiB")
