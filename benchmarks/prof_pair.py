import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lidar_slam_b200 as b2
from lidar_slam_b200 import synth
eng = b2.Engine(0)
src, tgt, T = synth.make_pair(0)
sd, _ = eng.voxel_downsample(eng.pack(src.numpy()), 0.02)
td, _ = eng.voxel_downsample(eng.pack(tgt.numpy()), 0.02)
p = b2.make_params(0.05, 12, rel_fitness=0.0, rel_rmse=0.0)
eng.set_icp_driver(sys.argv[1])
for _ in range(2):
    eng.icp(sd, td, p)
eng.synchronize()
