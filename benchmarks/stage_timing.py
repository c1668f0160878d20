#!/usr/bin/env python
"""Device time of every stage of the path on BASELINE-shaped inputs (CUDA events on the library stream)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lidar_slam_b200 as b2  # noqa: E402
from lidar_slam_b200 import synth  # noqa: E402


def timed(eng, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
    eng.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / reps


def main():
    eng = b2.Engine(0)
    src, tgt, T = synth.make_pair(0)
    s4, t4 = eng.pack(src.numpy()), eng.pack(tgt.numpy())
    sd, _ = eng.voxel_downsample(s4, 0.02)
    td, _ = eng.voxel_downsample(t4, 0.02)
    nrm = eng.estimate_normals(td, 0.04, 20)
    out = {"n_scan": int(s4.shape[0]), "n_down": int(sd.shape[0])}
    out["voxel_downsample_49k_2cm_us"] = round(timed(eng, lambda: eng.voxel_downsample(s4, 0.02, want_keys=False)), 1)
    out["estimate_normals_29k_k20_us"] = round(timed(eng, lambda: eng.estimate_normals(td, 0.04, 20)), 1)
    out["nn_search_29k_vs_29k_us"] = round(timed(eng, lambda: eng.nn_search(sd, td, 0.05)), 1)
    p30 = b2.make_params(0.05, 30, rel_fitness=0.0, rel_rmse=0.0)
    out["icp_p2p_pair_30it_us"] = round(timed(eng, lambda: eng.icp(sd, td, p30)), 1)
    p30l = b2.make_params(0.05, 30, mode=b2.P2L, rel_fitness=0.0, rel_rmse=0.0)
    out["icp_p2l_pair_30it_us"] = round(timed(eng, lambda: eng.icp(sd, td, p30l, tgt_normals=nrm)), 1)
    eng.set_icp_driver("sweep")  # the batch driver forced onto a single pair (auto picks the group kernel here)
    out["icp_p2p_pair_30it_sweep_us"] = round(timed(eng, lambda: eng.icp(sd, td, p30)), 1)
    out["icp_p2l_pair_30it_sweep_us"] = round(timed(eng, lambda: eng.icp(sd, td, p30l, tgt_normals=nrm)), 1)
    out["nn_search_29k_vs_29k_sweep_us"] = round(timed(eng, lambda: eng.nn_search(sd, td, 0.05)), 1)
    eng.set_icp_driver("auto")
    out["transform_49k_us"] = round(timed(eng, lambda: eng.transform(s4, T)), 1)
    eng.map_init(0.05, [-9.0, -9.0, -9.0], 0.05, 1 << 20)
    eng.map_fuse(t4)
    out["map_fuse_49k_5cm_us"] = round(timed(eng, lambda: eng.map_fuse(s4, T)), 1)
    rng = np.random.default_rng(0)
    cloud = eng.pack(np.concatenate([src.numpy(), tgt.numpy()]).astype(np.float32) + rng.normal(scale=1e-4, size=(98304, 3)).astype(np.float32))
    out["radius_outlier_98k_us"] = round(timed(eng, lambda: eng.radius_outlier_mask(cloud, 8, 0.023), reps=5), 1)
    out["statistical_outlier_98k_k45_us"] = round(timed(eng, lambda: eng.statistical_outlier_mask(cloud, 45, 3.5), reps=3, warm=1), 1)
    # row f3: PointCloud2 payload (device-resident bytes) -> float4 rows, NaN records dropped, projection fused
    k = synth.intrinsics()
    for n, tag in ((49152, "49k"), (2_000_000, "2M")):
        rec = np.stack([rng.integers(0, 256, n), rng.integers(0, 192, n), rng.uniform(0.3, 5.0, n)], axis=1).astype(np.float32)
        rec[::97, 2] = np.nan
        payload = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
        us = timed(eng, lambda: eng.ingest_pointcloud2(payload, n, 12, (0, 4, 8), intrinsics=k), reps=10)
        out[f"ingest_pointcloud2_{tag}_us"] = round(us, 1)
        out[f"ingest_pointcloud2_{tag}_GBps"] = round((12 * n + 16 * n + 16 * n + 16 * n) / us / 1e3, 1)  # payload + staged rows w/r + compacted rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
