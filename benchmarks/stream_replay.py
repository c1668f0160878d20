#!/usr/bin/env python
"""BASELINE config 4: streaming replay of a synthetic rosbag-shaped sequence through the plugin.

Every frame goes through `B200ICPProcessor.construct_global_map` (the reference's plugin entry point)
with HOST scans: upload -> 2 cm down-sample -> ICP against the growing resident map (constant-pose
prior, reference defaults: 5 cm gate, 1e-6/1e-6) -> fuse.  Prints one JSON line with the sustained
scans/s, the map size over time and the drift against the synthetic ground truth.

    python benchmarks/stream_replay.py --frames 1000 --map-voxel 0.02
"""
import argparse
import json
import multiprocessing
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--map-voxel", type=float, default=0.02)
    ap.add_argument("--max-iteration", type=int, default=30)
    ap.add_argument("--step", type=float, default=0.025)
    ap.add_argument("--capacity", type=int, default=1 << 22)
    args = ap.parse_args()

    import torch
    import lidar_slam_b200  # noqa: F401
    from lidar_slam_b200 import synth
    from lidar_slam_b200.plugin import DataTransfer, set_algorithm

    length = args.frames * args.step + 12.0
    scene = synth.make_corridor_scene(4242, length=length)
    poses = synth.corridor_trajectory(scene, args.frames, start_x=3.0, step=args.step, seed=7)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t0 = time.perf_counter()
    scans = [torch.from_numpy(synth.render_scan(scene, p, seed=k, device=dev).cpu().numpy()).pin_memory().numpy()
             for k, p in enumerate(poses)]
    t_gen = time.perf_counter() - t0

    cfg = os.path.join(ROOT, "3d-lidar-slam_b200", "config", "lidar_config.yaml")
    proc = set_algorithm("b200_icp", cfg, DataTransfer(), multiprocessing.Event(), mode="resident",
                         map_voxel=args.map_voxel, max_iteration=args.max_iteration,
                         map_capacity_cells=args.capacity)
    eng = proc._engine
    sizes, errs, fits, iters = [], [], [], []
    inv0 = np.linalg.inv(poses[0])
    t0 = time.perf_counter()
    for k, s in enumerate(scans):
        proc.construct_global_map(s)
        proc.point_clouds_in_map += 1
        if k and proc.last_result is not None:
            T_true = inv0 @ poses[k]
            errs.append(float(np.linalg.norm(proc.previous_transformation[-1][:3, 3] - T_true[:3, 3])))
            fits.append(proc.last_result.fitness)
            iters.append(proc.last_result.iterations)
        if k % max(1, args.frames // 10) == 0:
            sizes.append(eng.map_size())
    eng.synchronize()
    dt = time.perf_counter() - t0
    sizes.append(eng.map_size())
    print(json.dumps({
        "workload": f"config[3]: streaming replay, {args.frames} frames of 49,152 pts, map voxel {args.map_voxel} m",
        "scans_per_s_e2e": round(args.frames / dt, 1), "seconds": round(dt, 3), "scan_synthesis_s": round(t_gen, 1),
        "map_voxels_over_time": sizes, "final_map_voxels": sizes[-1],
        "drift_m": {"max": round(max(errs), 4), "final": round(errs[-1], 4)},
        "mean_fitness": round(float(np.mean(fits)), 4), "mean_iterations": round(float(np.mean(iters)), 2),
        "gpu_launches": eng.launch_count()}))


if __name__ == "__main__":
    main()
