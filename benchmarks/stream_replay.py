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
    ap.add_argument("--loop", action="store_true",
                    help="also replay the wire-format records through the queue + plugin.ProcessLoop (row f4)")
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
    loop_line = None
    if args.loop:
        import threading

        from lidar_slam_b200.plugin.process_pc_manager import ProcessLoop

        n_loop = min(args.frames, 400)
        recs = [synth.scan_records(synth.render_depth(scene, p, seed=k, device=dev)).cpu().numpy()
                for k, p in enumerate(poses[:n_loop])]
        dt2 = DataTransfer()
        proc2 = set_algorithm("b200_icp", cfg, dt2, multiprocessing.Event(), mode="resident",
                              map_voxel=args.map_voxel, max_iteration=args.max_iteration,
                              map_capacity_cells=args.capacity)
        stop, reset = multiprocessing.Event(), multiprocessing.Event()
        loop = ProcessLoop(proc2, dt2, stop, reset, max_batch=64, idle_timeout=0.05)
        th = threading.Thread(target=loop.run, daemon=True)
        t0 = time.perf_counter()
        th.start()
        for r in recs:
            dt2.pixel_depth_map_queue.put(r)
        while loop.scans_processed < n_loop and time.perf_counter() - t0 < 600:
            time.sleep(0.002)
        dtl = time.perf_counter() - t0
        stop.set()
        th.join(timeout=10)
        # nobody consumes the published maps here: do not let the queue's feeder thread block interpreter exit
        for q in (dt2.global_map_queue, dt2.pixel_depth_map_queue):
            q.cancel_join_thread()
        loop_line = {"frames": n_loop, "scans_per_s": round(n_loop / dtl, 1), "batches": loop.batches,
                     "maps_published": loop.maps_published,
                     "note": "records pickled through multiprocessing.Queue as in the reference; batch dequeue, "
                             "GPU projection, deferred pose read-back, map exported once per batch"}
    print(json.dumps({
        "process_loop": loop_line,
        "workload": f"config[3]: streaming replay, {args.frames} frames of 49,152 pts, map voxel {args.map_voxel} m",
        "scans_per_s_e2e": round(args.frames / dt, 1), "seconds": round(dt, 3), "scan_synthesis_s": round(t_gen, 1),
        "map_voxels_over_time": sizes, "final_map_voxels": sizes[-1],
        "drift_m": {"max": round(max(errs), 4), "final": round(errs[-1], 4)},
        "mean_fitness": round(float(np.mean(fits)), 4), "mean_iterations": round(float(np.mean(iters)), 2),
        "gpu_launches": eng.launch_count()}))


if __name__ == "__main__":
    main()
