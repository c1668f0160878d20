"""Developer script: device time per ICP step of the streaming path (49k-pt scan, 2 cm down-sampled, against the
2.1 M-voxel bench map), plus the whole scan step.  Run it under different B2_* environment switches to A/B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lidar_slam_b200 as b2
import bench

eng = b2.Engine(0)
dev = torch.device("cuda", 0)
scene = bench.build_scene(0)
nv = bench.preload_map(eng, scene, dev)
scans, poses = bench.make_scans(scene, 40, dev, seed=1)
sd, _ = eng.voxel_downsample(eng.pack(scans[1]), 0.02, want_keys=False)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("B2_"))
for iters in (30, 120):
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
    us = ev0.elapsed_time(ev1) * 1e3 / 20
    print(f"[{tag}] map_icp iters={r.iterations} fitness={r.fitness:.4f} rmse={r.inlier_rmse:.6f} "
          f"us/registration={us:.1f} us/step={us / (iters + 1):.2f}  T03={r.matrix()[0, 3]:.9f}")
# whole scan step, device-resident scans, graph replay
prm = b2.make_params(0.05, 30)
d4 = [eng.pack(s) for s in scans]
for rep in range(3):
    eng.set_pose(poses[0])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record()
        for s in d4:
            eng.slam_scan_dev(s, 0.02, prm)
        ev1.record()
    eng.synchronize()
    res = eng.slam_results(len(d4))
print(f"[{tag}] scan step: {ev0.elapsed_time(ev1) * 1e3 / len(d4):.1f} us/scan = {len(d4) / ev0.elapsed_time(ev1) * 1e3:.0f} scans/s, "
      f"mean iters {np.mean([x.iterations for x in res[1:]]):.1f}, final err "
      f"{np.linalg.norm(res[-1].matrix()[:3, 3] - poses[-1][:3, 3]):.4f} m, map {nv}")
