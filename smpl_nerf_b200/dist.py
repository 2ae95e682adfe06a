"""Multi-GPU frame rendering: rays shard, weights replicate, ONE all-gather of the rendered tiles.

The reference has no distributed code at all; rays are independent in every
pipeline (no cross-ray op), so the N-GPU path is a contiguous partition of the flattened
(image, row, col) ray index -- one process per GPU -- with no data-path collective, followed by a
single ``all_gather`` of the ``rgb_fine`` tiles (393 KB per rank for a 512x512 frame on 8 GPUs) so that
every rank holds the whole image for PSNR.  ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests) is the plumbing; the render itself is the fused kernel.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of rank's rays (the first n % world ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world {rank}/{world}')
    base, extra = divmod(n_rays, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_data(data: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    n = int(data[0].shape[0])
    a, b = shard_range(n, rank, world)
    return [t[a:b] for t in data]


def gather_tiles(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank row blocks [n_local, ...] of a contiguous partition into [n_total, ...].

    One collective: blocks are padded to the largest shard so a single ``all_gather_into_tensor``
    (NCCL) / ``all_gather`` (gloo) suffices; padding rows are dropped on assembly."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = math.ceil(n_total / world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    if local.is_cuda:
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        blocks = list(out.split(per, 0))
    else:
        blocks = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(blocks, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        parts.append(blocks[r][:b - a])
    return torch.cat(parts, 0)


def render_frame_sharded(render_fn: Callable[[List[torch.Tensor]], Sequence[torch.Tensor]],
                         data: Sequence[torch.Tensor], group=None, out_index: int = 1) -> torch.Tensor:
    """Render this rank's contiguous share of ``data``'s rays with ``render_fn`` (a pipeline's
    ``forward``) and return the assembled ``out[out_index]`` (``rgb_fine`` by default, as
    inference.py:252 reads it) for ALL rays, identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(data[0].shape[0])
    mine = shard_data(data, rank, world)
    local = render_fn(mine)[out_index]
    return gather_tiles(local.contiguous(), n, group)


def mse2psnr(mse: torch.Tensor) -> torch.Tensor:
    """-10 log10(mse): the formula of utils.py:484-488 / util/scores.py:47-48."""
    return -10.0 * torch.log10(mse)


def psnr(img: torch.Tensor, ref: torch.Tensor) -> float:
    return float(mse2psnr(torch.mean((img.double() - ref.double()) ** 2)))
