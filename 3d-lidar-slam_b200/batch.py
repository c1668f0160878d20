"""Sharding of independent scan pairs across ranks.

Single scan->map registration is sequential by construction (each ICP is seeded by the
previous pose and targets the map holding all earlier scans: icp_processor.py:106,118,74),
so the streaming path runs as independent replicas.  A batch of independent pairs
(loop-closure candidates, offline replay) shards with no communication during compute:
rank r owns the contiguous block ``shard_range(n_pairs, r, world)`` and the only
collectives are, at the end of the batch, one all-gather of the fixed-size result records
(152 B per pair) and — when every rank also fused its aligned scans into a partial map —
the gather of the partial maps as ``(voxel key, sum xyz, count)`` records followed by a
local count-weighted merge.  NCCL over NVLink on the GPUs, gloo in the CPU tests.
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


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


class ResultGatherer:
    """All-gather of the per-rank result records with every buffer allocated once (a gather per batch must
    not pay allocator calls, a zero-fill and a concatenation for a 152-byte-per-pair payload).

    ``gather(local)`` returns the global ``[n_items, itemsize]`` uint8 tensor, identical on every rank.  When the
    shards are equal (``n_items % world == 0``: config 5, 8192 pairs over 1/2/4/8 ranks) the collective writes
    straight into it; otherwise it runs on blocks padded to the largest shard and the ragged tail of every
    block is dropped with one ``index_select`` on a precomputed index."""

    def __init__(self, n_items: int, itemsize: int, device, group=None):
        import torch

        self.n_items, self.itemsize, self.group = int(n_items), int(itemsize), group
        dist = _dist()
        self.world = dist.get_world_size(group) if dist else 1
        self.rank = dist.get_rank(group) if dist else 0
        self.largest = -(-self.n_items // self.world)
        self.equal = self.n_items % self.world == 0
        self.out = torch.empty((self.world * self.largest, self.itemsize), dtype=torch.uint8, device=device)
        self.padded = None
        self.index = None
        if not self.equal:
            self.padded = torch.zeros((self.largest, self.itemsize), dtype=torch.uint8, device=device)
            idx = []
            for r in range(self.world):
                b, e = shard_range(self.n_items, r, self.world)
                idx.extend(range(r * self.largest, r * self.largest + (e - b)))
            self.index = torch.tensor(idx, dtype=torch.int64, device=device)

    def gather(self, local_records):
        dist = _dist()
        if dist is None or self.world == 1:
            return local_records
        if self.equal:
            dist.all_gather_into_tensor(self.out, local_records, group=self.group)
            return self.out
        self.padded[: local_records.shape[0]].copy_(local_records)
        dist.all_gather_into_tensor(self.out, self.padded, group=self.group)
        return self.out.index_select(0, self.index)


def gather_results(local_records, n_items: int, group=None):
    """One-shot form of :class:`ResultGatherer` (allocates its buffers per call)."""
    if _dist() is None:
        return local_records
    return ResultGatherer(n_items, local_records.shape[1], local_records.device, group).gather(local_records)


def records_view(records, engine=None) -> np.ndarray:
    """uint8 [n, itemsize] tensor -> structured numpy view (transformation, fitness, ...).  Pass the engine that
    produced the records when they may still be in flight on its stream."""
    if engine is not None:
        engine.synchronize()
    return records.cpu().numpy().view(RESULT_DTYPE).reshape(-1)


def gather_map_records(keys, sums, group=None):
    """All-gather every rank's partial map as records ``(voxel key int32[3], (sum x, sum y, sum z, count)
    float64[4])`` -> list over ranks of ``(keys [n_r,3], sums [n_r,4])`` views, identical on every rank.
    One small all-gather of the record counts, then one padded all-gather per array."""
    import torch

    dist = _dist()
    if dist is None:
        return [(keys, sums)]
    world = dist.get_world_size(group)
    n_local = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    counts = torch.zeros(world, dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = [int(v) for v in counts.cpu().tolist()]
    largest = max(max(counts), 1)
    pk = torch.zeros((largest, 3), dtype=torch.int32, device=keys.device)
    ps = torch.zeros((largest, 4), dtype=torch.float64, device=keys.device)
    pk[: keys.shape[0]] = keys
    ps[: sums.shape[0]] = sums
    ok = torch.empty((world * largest, 3), dtype=torch.int32, device=keys.device)
    os_ = torch.empty((world * largest, 4), dtype=torch.float64, device=keys.device)
    dist.all_gather_into_tensor(ok, pk, group=group)
    dist.all_gather_into_tensor(os_, ps, group=group)
    return [(ok[r * largest: r * largest + counts[r]], os_[r * largest: r * largest + counts[r]])
            for r in range(world)]


def merge_partial_maps(engine, group=None) -> int:
    """Replace this rank's resident map by the merge of ALL ranks' partial maps: export the local map as
    ``(key, sum xyz, count)`` records, all-gather them, reset the map and merge every rank's records in rank
    order with the count-weighted centroid merge (``b2_map_merge_records``).  Per voxel the sums and the counts
    add up, so the result equals ``voxel_down_sample`` of the concatenation of all ranks' points
    (icp_processor.py:74, 132-133) and is bit-identical on every rank.  Returns the merged voxel count."""
    import torch

    keys, sums = engine.map_export_records()
    with torch.cuda.stream(engine.stream):  # the collective is ordered after the export on the engine's stream
        parts = gather_map_records(keys, sums, group)
        engine.map_reset()
        for k, s in parts:
            if k.shape[0]:
                engine.map_merge_records(k.contiguous(), s.contiguous())
    return engine.map_size()
