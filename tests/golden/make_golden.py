"""Generates tests/golden/oracle_pins.json.

These are REGRESSION PINS OF THE ORACLE, not reference outputs: the reference publishes no golden vector
for this path and Open3D cannot run here (PARITY UNPINNED, see oracle/oracle.cpp).  They freeze what the
oracle computed on the seeded synthetic inputs when its known-answer tests and the NumPy/SciPy twin agreed,
so later edits of the oracle (or of the synthetic generator) cannot drift silently.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lidar_slam_b200  # noqa: E402,F401
from lidar_slam_b200 import synth  # noqa: E402
from oracle import cpu  # noqa: E402


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def main():
    src, tgt, T_true = synth.make_pair(0)
    src, tgt = src.numpy(), tgt.numpy()
    s64, t64 = src.astype(np.float64), tgt.astype(np.float64)
    sd, sk, sc = cpu.voxel_downsample(s64, 0.02)
    td, tk, tc = cpu.voxel_downsample(t64, 0.02)
    fused, fk, fc = cpu.voxel_downsample(np.concatenate([t64, cpu.transform(s64, T_true)]), 0.05,
                                         origin=[-9.0, -9.0, -9.0])
    icp = cpu.registration_icp(sd, td, 0.05, max_iteration=30)
    nrm = cpu.estimate_normals(td[:2000], 0.04, 20)
    idx, d2 = cpu.nn_search(sd, td, 0.05)
    out = {
        "inputs": {"src_sha": digest(src), "tgt_sha": digest(tgt), "n_src": int(len(src)), "n_tgt": int(len(tgt))},
        "voxel_0.02": {"n_src": int(len(sk)), "n_tgt": int(len(tk)), "src_keys_sha": digest(sk),
                       "src_counts_sha": digest(sc), "src_centroids_f32_sha": digest(sd.astype(np.float32))},
        "fuse_0.05": {"n": int(len(fk)), "keys_sha": digest(fk), "counts_sha": digest(fc)},
        "nn_0.05": {"n_corr": int((idx >= 0).sum()), "idx_sha": digest(idx)},
        "icp_p2p_30": {"T": icp.transformation.tolist(), "fitness": icp.fitness, "rmse": icp.inlier_rmse,
                       "iterations": icp.iterations, "corr_sha": digest(icp.correspondence)},
        "normals_first5": np.abs(nrm[:5]).tolist(),
    }
    path = os.path.join(ROOT, "tests", "golden", "oracle_pins.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
