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
