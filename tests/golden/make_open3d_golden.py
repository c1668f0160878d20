"""Regenerates the golden pins from REAL Open3D calls — the route from "parity unpinned" to pinned parity.

Runs only where ``import open3d`` works (it does not in the build container: Open3D is an un-vendored conda
dependency of the reference, environment.yml:14).  On such a machine:

    python tests/golden/make_open3d_golden.py        # writes tests/golden/open3d_pins.json

and commit the file.  ``tests/test_golden.py::test_oracle_against_open3d_pins`` (skipped while the file is
absent) then checks the CPU oracle against it on the same seeded inputs, and — through the existing GPU-vs-oracle
suite — the CUDA path inherits the pin.  Every pin is produced by the very Open3D entry point the reference calls
(src/slam/scripts/point_cloud_processors/icp_processor.py:90-113, 130-146):

    voxel_down_sample(0.02)                                   icp_processor.py:90-91
    voxel_down_sample at the fused cloud, 0.05                icp_processor.py:132-133 (same call, other size)
    estimate_normals(KDTreeSearchParamHybrid(0.04, 20))       icp_processor.py:94-99
    registration_icp(..., 0.05, init, PointToPoint, (30, 1e-6, 1e-6))   icp_processor.py:103-113
    registration_icp(..., PointToPlane)                       (BASELINE config 3)
    remove_statistical_outlier(45, 3.5) / (80, 2.0), remove_radius_outlier(8, 0.023)   icp_processor.py:136-146

Open3D returns down-sampled points in hash-map order and does not report the ICP iteration count; both are
canonicalised here: rows are sorted by voxel key (Open3D's own origin, min_bound - voxel/2), and the iteration
count is the smallest budget whose result equals the 30-iteration result.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def to_pcd(o3d, pts):
    p = o3d.geometry.PointCloud()
    p.points = o3d.utility.Vector3dVector(np.asarray(pts, dtype=np.float64))  # icp_processor.py:60,65
    return p


def canonical_downsample(o3d, pts, voxel, origin=None):
    """voxel_down_sample + rows sorted by voxel key -> (centroids f64 [m,3], keys int32 [m,3])."""
    pts = np.asarray(pts, dtype=np.float64)
    if origin is None:
        down = np.asarray(to_pcd(o3d, pts).voxel_down_sample(voxel).points)
        org = pts.min(axis=0) - voxel * 0.5
    else:
        # Open3D has no explicit-origin overload: voxel_down_sample_and_trace takes the bounds it derives
        # the origin from (origin = min_bound - voxel/2)
        org = np.asarray(origin, dtype=np.float64)
        mn = org + voxel * 0.5
        down, _, _ = to_pcd(o3d, pts).voxel_down_sample_and_trace(voxel, mn, pts.max(axis=0) + voxel, False)
        down = np.asarray(down.points)
    keys = np.floor((down - org) / voxel).astype(np.int32)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return down[order], keys[order]


def icp(o3d, src, tgt, max_corr, init, max_iteration, plane=False):
    reg = o3d.pipelines.registration
    est = reg.TransformationEstimationPointToPlane() if plane else reg.TransformationEstimationPointToPoint()
    crit = reg.ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=max_iteration)
    return reg.registration_icp(src, tgt, max_corr, init, est, crit)


def corr_array(result, n_src):
    c = np.full(n_src, -1, dtype=np.int32)
    cs = np.asarray(result.correspondence_set)
    if len(cs):
        c[cs[:, 0]] = cs[:, 1]
    return c


def main():
    import open3d as o3d

    import lidar_slam_b200  # noqa: F401
    from lidar_slam_b200 import synth

    src, tgt, T_true = synth.make_pair(0)
    src, tgt = src.numpy(), tgt.numpy()
    s64, t64 = src.astype(np.float64), tgt.astype(np.float64)
    sd, sk = canonical_downsample(o3d, s64, 0.02)
    td, tk = canonical_downsample(o3d, t64, 0.02)
    moved = np.asarray(to_pcd(o3d, s64).transform(T_true).points)  # icp_processor.py:73
    fused, fk = canonical_downsample(o3d, np.concatenate([t64, moved]), 0.05, origin=[-9.0, -9.0, -9.0])

    psd, ptd = to_pcd(o3d, sd), to_pcd(o3d, td)
    tree = o3d.geometry.KDTreeFlann(ptd)
    idx = np.full(len(sd), -1, dtype=np.int32)
    for i, q in enumerate(sd):                      # A.3 Eval: SearchHybrid(q, max_corr, 1)
        k, j, _ = tree.search_hybrid_vector_3d(q, 0.05, 1)
        if k:
            idx[i] = j[0]

    full = icp(o3d, psd, ptd, 0.05, np.eye(4), 30)
    iters = 30
    for k in range(0, 31):
        r = icp(o3d, psd, ptd, 0.05, np.eye(4), k)
        if np.array_equal(r.transformation, full.transformation) and r.fitness == full.fitness:
            iters = k
            break

    ptd_n = to_pcd(o3d, td[:2000])
    ptd_n.estimate_normals(o3d.geometry.KDTreeSearchParamHybrid(radius=0.04, max_nn=20))
    nrm = np.asarray(ptd_n.normals)

    ptd_full = to_pcd(o3d, td)
    ptd_full.estimate_normals(o3d.geometry.KDTreeSearchParamHybrid(radius=0.04, max_nn=20))
    p2l = icp(o3d, psd, ptd_full, 0.05, np.eye(4), 30, plane=True)

    cloud = to_pcd(o3d, np.concatenate([td, sd]))
    _, keep_a = cloud.remove_statistical_outlier(nb_neighbors=45, std_ratio=3.5)
    _, keep_b = cloud.remove_statistical_outlier(nb_neighbors=80, std_ratio=2.0)
    _, keep_c = cloud.remove_radius_outlier(nb_points=8, radius=0.023)

    def mask(ind, n):
        m = np.zeros(n, dtype=np.uint8)
        m[np.asarray(ind, dtype=np.int64)] = 1
        return m

    n_cloud = len(td) + len(sd)
    out = {
        "source": f"open3d {o3d.__version__}",
        "inputs": {"src_sha": digest(src), "tgt_sha": digest(tgt), "n_src": int(len(src)), "n_tgt": int(len(tgt))},
        "voxel_0.02": {"n_src": int(len(sk)), "n_tgt": int(len(tk)), "src_keys_sha": digest(sk),
                       "src_centroids_f32_sha": digest(sd.astype(np.float32)),
                       "src_centroids_first5": sd[:5].tolist()},
        "fuse_0.05": {"n": int(len(fk)), "keys_sha": digest(fk), "centroids_first5": fused[:5].tolist()},
        "nn_0.05": {"n_corr": int((idx >= 0).sum()), "idx_sha": digest(idx)},
        "icp_p2p_30": {"T": np.asarray(full.transformation).tolist(), "fitness": full.fitness,
                       "rmse": full.inlier_rmse, "iterations": iters,
                       "corr_sha": digest(corr_array(full, len(sd)))},
        "icp_p2l_30": {"T": np.asarray(p2l.transformation).tolist(), "fitness": p2l.fitness, "rmse": p2l.inlier_rmse},
        "normals_first5": np.abs(nrm[:5]).tolist(),
        "sor_45_3.5": {"kept": int(len(keep_a)), "mask_sha": digest(mask(keep_a, n_cloud))},
        "sor_80_2.0": {"kept": int(len(keep_b)), "mask_sha": digest(mask(keep_b, n_cloud))},
        "ror_8_0.023": {"kept": int(len(keep_c)), "mask_sha": digest(mask(keep_c, n_cloud))},
    }
    path = os.path.join(ROOT, "tests", "golden", "open3d_pins.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    try:
        import open3d  # noqa: F401
    except ImportError:
        sys.exit("open3d is not importable here: run this script where the reference's environment.yml is installed")
    main()
