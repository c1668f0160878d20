"""world_size-2 gloo test of the batch sharding + result gather (the N>1 host path, on CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    import lidar_slam_b200  # noqa: F401
    from lidar_slam_b200._lib import RESULT_DTYPE
    from lidar_slam_b200.batch import gather_results, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    b, e = shard_range(n_pairs, rank, world)
    rec = np.zeros(e - b, dtype=RESULT_DTYPE)
    for i, pair in enumerate(range(b, e)):  # stand-in for the per-rank GPU batch: results keyed by pair id
        rec["transformation"][i] = np.eye(4) * (pair + 1)
        rec["fitness"][i] = pair / n_pairs
        rec["iterations"][i] = pair
    local = torch.from_numpy(rec.view(np.uint8).reshape(e - b, RESULT_DTYPE.itemsize).copy())
    allrec = gather_results(local, n_pairs)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allrec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_every_pair(tmp_path):
    from lidar_slam_b200._lib import RESULT_DTYPE

    n_pairs, world = 11, 2  # odd count: shards differ in size
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(outs[0], outs[1])
    rec = outs[0].view(RESULT_DTYPE).reshape(-1)
    assert len(rec) == n_pairs
    np.testing.assert_array_equal(rec["iterations"], np.arange(n_pairs))
    for pair in range(n_pairs):
        np.testing.assert_array_equal(rec["transformation"][pair], np.eye(4) * (pair + 1))


def _map_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import lidar_slam_b200  # noqa: F401
    from lidar_slam_b200.batch import ResultGatherer, gather_map_records

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # rank r holds a partial map of 3 + 2r voxels (ragged), records keyed by (rank, i)
    n = 3 + 2 * rank
    keys = torch.tensor([[rank, i, -i] for i in range(n)], dtype=torch.int32)
    sums = torch.tensor([[rank + 0.5, i, 2.0 * i, i + 1] for i in range(n)], dtype=torch.float64)
    parts = gather_map_records(keys, sums)
    np.savez(os.path.join(out_dir, f"map{rank}.npz"), **{f"k{r}": k.numpy() for r, (k, _) in enumerate(parts)},
             **{f"s{r}": s.numpy() for r, (_, s) in enumerate(parts)})
    # pre-allocated gatherer: equal shards (no padding) and ragged shards, called twice (buffers are reused)
    for n_items in (8, 7):
        g = ResultGatherer(n_items, 16, "cpu")
        b = n_items // world * rank + min(rank, n_items % world)
        e = b + n_items // world + (1 if rank < n_items % world else 0)
        for rep in range(2):
            local = torch.arange(b, e, dtype=torch.uint8).reshape(-1, 1).repeat(1, 16) + rep
            allrec = g.gather(local)
            assert allrec.shape == (n_items, 16)
            assert torch.equal(allrec[:, 0], torch.arange(n_items, dtype=torch.uint8) + rep)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partial_map_gather_carries_keys_sums_and_counts(tmp_path):
    world = 2
    port = 31000 + os.getpid() % 2000
    mp.spawn(_map_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"map{r}.npz") for r in range(world)]
    for r in range(world):
        n = 3 + 2 * r
        for o in outs:  # identical on every rank
            np.testing.assert_array_equal(o[f"k{r}"], np.array([[r, i, -i] for i in range(n)], np.int32))
            np.testing.assert_array_equal(o[f"s{r}"], np.array([[r + 0.5, i, 2.0 * i, i + 1] for i in range(n)]))
