"""Sharding of independent scan pairs across ranks.

Single scan->map registration is sequential by construction (each ICP is seeded by the
previous pose and targets the map holding all earlier scans: icp_processor.py:106,118,74),
so the streaming path runs as independent replicas.  A batch of independent pairs
(loop-closure candidates, offline replay) shards with no communication during compute:
rank r owns the contiguous block ``shard_range(n_pairs, r, world)`` and the only
collective is one all-gather of the fixed-size result records (152 B per pair) at the
end of the batch — NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from ._lib import RESULT_DTYPE


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition; the first ``n_items % world`` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_results(local_records, n_items: int, group=None):
    """All-gather per-rank result records (uint8 tensor [n_local, itemsize]) into the global
    [n_items, itemsize] tensor, identical on every rank.  Shards may differ by one item, so the
    collective runs on blocks padded to the largest shard."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_records
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    item = local_records.shape[1]
    largest = -(-n_items // world)
    padded = torch.zeros((largest, item), dtype=torch.uint8, device=local_records.device)
    padded[: local_records.shape[0]] = local_records
    out = torch.empty((world * largest, item), dtype=torch.uint8, device=local_records.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(n_items, r, world)
        parts.append(out[r * largest: r * largest + (e - b)])
    del rank
    return torch.cat(parts, dim=0)


def records_view(records) -> np.ndarray:
    """uint8 [n, itemsize] tensor -> structured numpy view (transformation, fitness, ...)."""
    return records.cpu().numpy().view(RESULT_DTYPE).reshape(-1)


def merge_partial_maps(engine, local_pts, group=None):
    """Gather every rank's partial map (voxel centroids, float32 [n,4]) and merge them into this
    rank's resident map with the same centroid-merge kernel the streaming path uses.  Every gathered
    centroid enters the merge as one point (the per-voxel point counts are not re-weighted)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        engine.map_fuse(local_pts)
        return
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_pts.shape[0]], dtype=torch.int64, device=local_pts.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    largest = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros((largest, 4), dtype=torch.float32, device=local_pts.device)
    padded[: local_pts.shape[0]] = local_pts
    out = torch.empty((world * largest, 4), dtype=torch.float32, device=local_pts.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    for r in range(world):
        n = int(sizes[r].item())
        if n:
            engine.map_fuse(out[r * largest: r * largest + n].contiguous())
