"""Developer script: per-registration device time of map ICP (persistent vs per-launch path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lidar_slam_b200 as b2
import bench
eng = b2.Engine(0)
scene = bench.build_scene(0)
print("map voxels", bench.preload_map(eng, scene, torch.device("cuda", 0)))
scans, poses = bench.make_scans(scene, 3, torch.device("cuda", 0), seed=1)
sd, _ = eng.voxel_downsample(eng.pack(scans[1]), 0.02, want_keys=False)
for iters in (12, 60):
    prm = b2.make_params(0.05, iters, rel_fitness=0.0, rel_rmse=0.0)
    for _ in range(3):
        r = eng.map_icp(sd, prm, init=poses[1])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record()
        for _ in range(20):
            r = eng.map_icp(sd, prm, init=poses[1])
        ev1.record()
    eng.synchronize()
    print(f"iters={r.iterations} fitness={r.fitness:.4f} rmse={r.inlier_rmse:.6f} "
          f"us/registration={ev0.elapsed_time(ev1) * 1e3 / 20:.1f} us/step={ev0.elapsed_time(ev1) * 1e3 / 20 / (iters + 1):.2f}")
    print(np.array2string(r.matrix(), precision=6))
