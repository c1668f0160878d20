"""Profiling driver: every primitive of the path once on BASELINE-shaped single-scan inputs (run under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lidar_slam_b200 as b2  # noqa: E402
from lidar_slam_b200 import synth  # noqa: E402

eng = b2.Engine(0)
src, tgt, T = synth.make_pair(0)
s4, t4 = eng.pack(src.numpy()), eng.pack(tgt.numpy())
rng = np.random.default_rng(0)
cloud = eng.pack(np.concatenate([src.numpy(), tgt.numpy()]).astype(np.float32)
                 + rng.normal(scale=1e-4, size=(98304, 3)).astype(np.float32))
rec = np.stack([rng.integers(0, 256, 49152), rng.integers(0, 192, 49152), rng.uniform(0.3, 5.0, 49152)], axis=1).astype(np.float32)
payload = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
for rep in range(2):  # the second repetition is the one to read (workspaces sized, caches as in steady state)
    sd, _ = eng.voxel_downsample(s4, 0.02)
    td, _ = eng.voxel_downsample(t4, 0.02)
    nrm = eng.estimate_normals(td, 0.04, 20)
    eng.nn_search(sd, td, 0.05)
    eng.icp(sd, td, b2.make_params(0.05, 2))
    eng.icp(sd, td, b2.make_params(0.05, 2, mode=b2.P2L), tgt_normals=nrm)
    eng.transform(s4, T)
    eng.map_init(0.05, [-9.0, -9.0, -9.0], 0.05, 1 << 20)
    eng.map_fuse(t4)
    eng.map_fuse(s4, T)
    eng.map_icp(sd, b2.make_params(0.05, 2))
    eng.radius_outlier_mask(cloud, 8, 0.023)
    eng.statistical_outlier_mask(cloud, 45, 3.5)
    eng.ingest_pointcloud2(payload, 49152, 12, (0, 4, 8), intrinsics=synth.intrinsics())
    eng.synchronize()
print("done")
