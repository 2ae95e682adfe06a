"""World-size-2 (and 3) gloo tests of the multi-GPU host logic (smpl_nerf_b200/dist.py) on CPU.

The render itself needs a GPU, so ``render_fn`` here is a deterministic per-ray stand-in with the pipelines'
output convention (tuple, index 1 = rgb_fine[B,3]); what is under test is the contiguous ray partition,
the single padded all-gather and the assembly order -- every rank must end up with the image a single
process would have rendered."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smpl_nerf_b200 import dist as nd


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_pipeline(data):
    samples, origin, direction, z = data[:4]
    rgb = torch.stack([samples.sum((1, 2)), origin.sum(1) * z.mean(1), direction[:, 0] - z[:, -1]], -1)
    return rgb * 0.5, rgb, samples, z


def _make_data(n_rays: int, n_coarse: int = 8):
    g = torch.Generator().manual_seed(1234)
    return [torch.randn(n_rays, n_coarse, 3, generator=g), torch.randn(n_rays, 3, generator=g),
            torch.randn(n_rays, 3, generator=g), torch.rand(n_rays, n_coarse, generator=g).sort(-1).values,
            torch.rand(n_rays, 3, generator=g)]


def _worker(rank: int, world: int, port: int, n_rays: int, out_dir: str):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        data = _make_data(n_rays)
        img = nd.render_frame_sharded(_fake_pipeline, data)
        want = _fake_pipeline(data)[1]
        assert img.shape == want.shape
        assert torch.equal(img, want), f'rank {rank}: assembled image differs'
        a, b = nd.shard_range(n_rays, rank, world)
        # gather of a non-default output (alpha-like [n, k] block) keeps row order too
        z = nd.gather_tiles(data[3][a:b].contiguous(), n_rays)
        assert torch.equal(z, data[3])
        # whole-image PSNR is identical on every rank
        p = torch.tensor([nd.psnr(img, data[4])], dtype=torch.float64)
        ps = [torch.zeros_like(p) for _ in range(world)]
        dist.all_gather(ps, p)
        assert all(torch.equal(ps[0], q) for q in ps)
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_rays', [(2, 64), (2, 37), (3, 10), (2, 1)])
def test_sharded_frame_matches_single_process(tmp_path, world, n_rays):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_rays, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16384, 262144):
        for world in (1, 2, 3, 8):
            spans = [nd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        nd.shard_range(10, 2, 2)


def test_single_process_is_a_no_op():
    data = _make_data(9)
    img = nd.render_frame_sharded(_fake_pipeline, data)
    assert torch.equal(img, _fake_pipeline(data)[1])
    assert abs(nd.psnr(torch.zeros(4, 3), torch.full((4, 3), 0.1)) - 20.0) < 1e-4


def _warm_worker(rank: int, world: int, port: int, out_dir: str):
    """bench.py's warm-up: ranks whose steps take different host time must still run the SAME number of steps (a step holds a
    collective; a purely time-based loop deadlocked the 8-GPU run of round 2)."""
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        bench.WARM_SECONDS = 0.3
        dev = torch.device('cpu')
        t0, i = time.perf_counter(), 0
        while not bench.warm_done(i, t0, 3, world, dev):
            for _ in range(4):
                time.sleep(0.004 * (1 + 5 * rank))              # rank 1 is 6x slower per step
                x = torch.ones(1)
                dist.all_reduce(x)                               # the collective inside a step: a count mismatch would hang here
                i += 1
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([i], dtype=torch.int64))
        assert all(int(c) == i for c in counts) and i >= 3
        open(os.path.join(out_dir, f'warm{rank}'), 'w').write(str(i))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bench_warm_up_runs_the_same_number_of_steps_on_every_rank(tmp_path):
    port = _free_port()
    mp.spawn(_warm_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'warm{r}').exists() for r in range(2))
