"""Developer script: the scan step with the A.3 loop as chained launches vs as a CUDA-graph WHILE node
(B2_ICP_WHILE=0/1), for a short (30) and the reference's default (500) iteration budget."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lidar_slam_b200 as b2
import bench
eng = b2.Engine(0)
dev = torch.device("cuda", 0)
scene = bench.build_scene(0)
bench.preload_map(eng, scene, dev)
scans, poses = bench.make_scans(scene, 48, dev, seed=1)
d4 = [eng.pack(s) for s in scans]
for iters in (30, 500):
    prm = b2.make_params(0.05, iters)
    for rep in range(2):
        eng.set_pose(poses[0])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(eng.stream):
            ev0.record()
            for s in d4:
                eng.slam_scan_dev(s, 0.02, prm)
            ev1.record()
        eng.synchronize()
        res = eng.slam_results(len(d4))
    h = hashlib.sha1(np.concatenate([r.matrix().ravel() for r in res]).tobytes()).hexdigest()[:12]
    print(f"max_iter={iters} while={os.environ.get('B2_ICP_WHILE', 'auto')} us/scan={ev0.elapsed_time(ev1) * 1e3 / len(d4):.1f} "
          f"mean_iters={np.mean([r.iterations for r in res]):.1f} poses={h}")
