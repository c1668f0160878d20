"""Developer smoke script (not collected by pytest): stage-by-stage GPU-vs-oracle comparison with timings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
sys.path.insert(0, ROOT)
import lidar_slam_b200 as b2  # noqa: E402
from lidar_slam_b200 import synth  # noqa: E402
from oracle import cpu as orc  # noqa: E402


def main():
    eng = b2.Engine(0)
    src, tgt, T_true = synth.make_pair(0)
    src_np, tgt_np = src.numpy(), tgt.numpy()
    s4, t4 = eng.pack(src_np), eng.pack(tgt_np)

    # --- voxel downsample
    t0 = time.time()
    sd, sk = eng.voxel_downsample(s4, 0.02)
    eng.synchronize()
    print("gpu downsample", time.time() - t0, sd.shape)
    osd, osk, osc = orc.voxel_downsample(src_np.astype(np.float64), 0.02)
    print("keys equal:", np.array_equal(sk.cpu().numpy(), osk), "n", len(osk))
    g = sd.cpu().numpy()
    print("centroid bit-exact (f32):", np.array_equal(g[:, :3], osd.astype(np.float32)),
          "counts equal:", np.array_equal(g[:, 3].view(np.int32), osc))
    td, tk = eng.voxel_downsample(t4, 0.02)
    otd, _, _ = orc.voxel_downsample(tgt_np.astype(np.float64), 0.02)

    # --- nn search, teacher forced (oracle sees the same f32 clouds)
    T0 = np.eye(4)
    T0[:3, 3] = [0.01, -0.005, 0.003]
    idx, d2 = eng.nn_search(sd, td, 0.05, T=T0)
    sdf = sd.cpu().numpy()[:, :3].astype(np.float64)
    tdf = td.cpu().numpy()[:, :3].astype(np.float64)
    oidx, od2 = orc.nn_search(orc.transform(sdf, T0), tdf, 0.05)
    gi = idx.cpu().numpy()
    print("nn idx equal:", np.array_equal(gi, oidx), "mismatch", int((gi != oidx).sum()),
          "d2 bit-equal:", np.array_equal(d2.cpu().numpy()[oidx >= 0], od2[oidx >= 0]))

    # --- full ICP
    prm = b2.make_params(0.05, 30)
    t0 = time.time()
    res, corr = eng.icp(sd, td, prm, want_corr=True)
    print("gpu icp wall", time.time() - t0, res.fitness, res.inlier_rmse, res.iterations)
    t0 = time.time()
    ores = orc.registration_icp(sdf, tdf, 0.05, max_iteration=30)
    print("cpu icp wall", time.time() - t0, ores.fitness, ores.inlier_rmse, ores.iterations)
    print("T maxdiff", np.abs(res.matrix() - ores.transformation).max(),
          "corr agree", float((corr.cpu().numpy() == ores.correspondence).mean()))

    # --- normals + p2l
    nrm = eng.estimate_normals(td, 0.04, 20)
    onrm = orc.estimate_normals(tdf, 0.04, 20)
    gn = nrm.cpu().numpy()[:, :3].astype(np.float64)
    cosang = np.abs((gn * onrm).sum(1))
    print("normals: frac |cos|>1-1e-8:", float((cosang > 1 - 1e-8).mean()), "min", cosang.min())
    prm2 = b2.make_params(0.05, 30, mode=b2.P2L)
    res2, _ = eng.icp(sd, td, prm2, tgt_normals=nrm)
    ores2 = orc.registration_icp(sdf, tdf, 0.05, max_iteration=30, mode="p2l", tgt_normals=gn)
    print("p2l T maxdiff", np.abs(res2.matrix() - ores2.transformation).max(), res2.iterations, ores2.iterations,
          res2.fitness, ores2.fitness)

    # --- resident map
    org = np.array([-10.0, -10.0, -10.0])
    eng.map_init(0.05, org, 0.05, 1 << 18)
    eng.map_fuse(t4, None)
    eng.synchronize()
    mp, mk = eng.map_export()
    omp, omk, omc = orc.voxel_downsample(tgt_np.astype(np.float64), 0.05, origin=org)
    print("map keys equal:", np.array_equal(mk.cpu().numpy(), omk), eng.map_size(), len(omk),
          "cent equal:", np.array_equal(mp.cpu().numpy()[:, :3], omp.astype(np.float32)))
    r3 = eng.map_icp(sd, prm)
    o3 = orc.registration_icp(sdf, omp.astype(np.float32).astype(np.float64), 0.05, max_iteration=30)
    print("map icp T maxdiff", np.abs(r3.matrix() - o3.transformation).max(), r3.iterations, o3.iterations)

    # --- slam scan path (brick 3)
    eng.map_init(0.02, org, 0.05, 1 << 18)
    pin = torch.from_numpy(tgt_np).pin_memory().numpy()
    r = eng.slam_scan(pin, 0.02, prm)
    pin2 = torch.from_numpy(src_np).pin_memory().numpy()
    t0 = time.time()
    r = eng.slam_scan(pin2, 0.02, prm)
    print("slam scan wall", time.time() - t0, r.fitness, r.inlier_rmse, r.iterations, "map", eng.map_size())
    ot, otk, _ = orc.voxel_downsample(tgt_np.astype(np.float64), 0.02, origin=org)
    o4 = orc.registration_icp(sdf, ot.astype(np.float32).astype(np.float64), 0.05, max_iteration=30)
    print("slam T maxdiff", np.abs(r.matrix() - o4.transformation).max(), r.iterations, o4.iterations)
    for k in range(5):
        t0 = time.time()
        eng.slam_scan(pin2, 0.02, prm)
        print("slam scan wall", time.time() - t0)
    print("launches", eng.launch_count())


if __name__ == "__main__":
    main()
