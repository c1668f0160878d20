"""Developer script: a longer chain of dense-map ICP steps (for ncu source-level sampling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lidar_slam_b200 as b2
import bench

eng = b2.Engine(0)
dev = torch.device("cuda", 0)
scene = bench.build_scene(0)
bench.preload_map(eng, scene, dev)
scans, poses = bench.make_scans(scene, 3, dev, seed=1)
sd, _ = eng.voxel_downsample(eng.pack(scans[1]), 0.02, want_keys=False)
prm = b2.make_params(0.05, 40, rel_fitness=0.0, rel_rmse=0.0)
for _ in range(2):
    eng.map_icp(sd, prm, init=poses[1])
