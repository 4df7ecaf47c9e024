"""Frame pairs of a clip sharded across the GPUs of one box.

Every (source, target) pair is independent (ofgen_keyframe_inpaint.py:585-600 walks a flat pair
list; the per-frame loop ofgen_pixel_inpaint.py:324-356 depends only on the key frame), so ranks
take contiguous ranges of the pair list -- which keeps pairs sharing a key frame on one rank --
and there is NO collective on the data path.  The only exchange is the optional final gather of
the [pairs, H, W, 3] flow+confidence stack (NCCL over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, end) of `n_items` for `rank`; sizes differ by at most one, earlier ranks
    take the extra items, empty ranges are allowed when n_items < world_size."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank/world_size {rank}/{world_size}')
    if n_items < 0:
        raise ValueError('n_items must be >= 0')
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_pairs(pairs: Sequence, rank: int, world_size: int) -> List:
    s, e = shard_range(len(pairs), rank, world_size)
    return list(pairs[s:e])


def shard_by_target(pairs: Sequence[Tuple[int, int]], rank: int, world_size: int) -> List[Tuple[int, int]]:
    """Sharding for the greedy composite (M5), which needs every reference of one TARGET frame on
    one rank: targets (in first-appearance order) are split contiguously, pairs follow their target."""
    targets: List[int] = []
    seen = set()
    for _, t in pairs:
        if t not in seen:
            seen.add(t)
            targets.append(t)
    s, e = shard_range(len(targets), rank, world_size)
    mine = set(targets[s:e])
    return [p for p in pairs if p[1] in mine]


def key_frame_pairs(n_frames: int, key_every: int) -> List[Tuple[int, int]]:
    """(key, frame) pairs of a clip with a key frame every `key_every` frames (config 5)."""
    return [((f // key_every) * key_every, f) for f in range(n_frames) if f % key_every != 0]


def gather_stack(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Reassemble the per-rank [n_local, ...] stacks (contiguous shard_range order) into [n_total, ...]
    on every rank.  Ranks may hold different n_local; one all_gather of padded slabs."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    slab = local.new_zeros((per, *local.shape[1:]))
    slab[: local.shape[0]] = local
    out = local.new_empty((world * per, *local.shape[1:]))
    dist.all_gather_into_tensor(out, slab, group=group)
    parts = []
    for r in range(world):
        s, e = shard_range(n_total, r, world)
        parts.append(out[r * per: r * per + (e - s)])
    return torch.cat(parts, dim=0)
