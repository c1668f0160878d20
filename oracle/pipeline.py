"""The reference's per-scan sequence on the CPU oracle.  TEST INFRASTRUCTURE.

A line-by-line restatement of ``ICPProcessor`` (src/slam/scripts/point_cloud_processors/
icp_processor.py:43-148) with every Open3D call replaced by the float64 oracle
(oracle/cpu.py).  PARITY UNPINNED (see oracle.cpp): neither Open3D nor any golden
vector of the reference is available.

``ReferenceSequence``  — the reference-faithful O(map)-per-scan sequence (raw map,
                         re-voxelise + re-index both clouds every scan).
``ResidentSequence``   — the same arithmetic with the map kept as voxel accumulators at
                         a fixed origin (what the GPU's resident mode computes); used
                         to check the resident path and as the "resident CPU variant"
                         baseline of BASELINE.md.
"""
from __future__ import annotations

import numpy as np

from . import cpu


class ReferenceSequence:
    def __init__(self, voxel_size=0.02, max_iteration=500, relative_fitness=1e-6, relative_rmse=1e-6,
                 maintenance="full", estimate_unused_normals=False, f32_storage=False):
        # f32_storage: round every stored cloud (down-sampled clouds, transformed scan) to float32,
        # which is how the GPU keeps points in HBM — the teacher-forced variant for parity tests.
        self.f32_storage = f32_storage
        self.voxel_size = voxel_size
        self.max_iteration = max_iteration
        self.relative_fitness = relative_fitness
        self.relative_rmse = relative_rmse
        self.maintenance = maintenance  # "full" | "voxel" | "off"
        self.estimate_unused_normals = estimate_unused_normals  # icp_processor.py:94-99 (result unused)
        self.global_map = None
        self.point_clouds_in_map = 0
        self.previous_transformation = [np.identity(4)]
        self.last_result = None

    def _store(self, a):
        return a.astype(np.float32).astype(np.float64) if self.f32_storage else a

    def process(self, points, queue_empty=False):
        """generic_point_cloud_processor.py:111-126 minus the queue and the projection."""
        self.construct_global_map(points, queue_empty)
        self.point_clouds_in_map += 1

    def construct_global_map(self, points, queue_empty=False):
        if self.point_clouds_in_map == 0:  # icp_processor.py:53-56
            self.point_clouds_in_map += 1
            self.global_map = np.asarray(points, dtype=np.float64)
            return
        scan = np.asarray(points, dtype=np.float64)  # Vector3dVector widening, icp_processor.py:60
        T = self.align_point_clouds_with_icp(scan, self.global_map)
        moved = self._store(cpu.transform(scan, T))  # icp_processor.py:73
        self.global_map = np.concatenate([moved, self.global_map], axis=0)  # scan rows first, :74
        self.do_stuff(queue_empty)

    def align_point_clouds_with_icp(self, source, target):
        v = self.voxel_size
        source_down = self._store(cpu.voxel_downsample(source, v)[0])  # :90
        target_down = self._store(cpu.voxel_downsample(target, v)[0])  # :91
        if self.estimate_unused_normals:
            cpu.estimate_normals(source_down, v * 2, 20)  # :94-96
            cpu.estimate_normals(target_down, v * 2, 20)  # :97-99
        res = cpu.registration_icp(source_down, target_down, v * 2.5, init=self.previous_transformation[-1],
                                   mode="p2p", max_iteration=self.max_iteration,
                                   relative_fitness=self.relative_fitness, relative_rmse=self.relative_rmse)
        self.last_result = res
        self.previous_transformation.append(res.transformation)  # :118
        return res.transformation

    def do_stuff(self, queue_empty=False):
        if self.maintenance == "off" or self.point_clouds_in_map < 10:  # :127-128
            return
        if self.point_clouds_in_map % 10 == 0 or queue_empty:  # :129
            m, _, _ = cpu.voxel_downsample(self.global_map, 0.001)  # :130-133
            if self.maintenance == "voxel":
                self.global_map = m
                return
            if self.point_clouds_in_map % 40 == 0:  # :135-143
                m = m[cpu.statistical_outlier_mask(m, 80, 2.0)]
                m = m[cpu.radius_outlier_mask(m, 8, 0.023)]
                self.global_map = m
                return
            self.global_map = m[cpu.statistical_outlier_mask(m, 45, 3.5)]  # :145-148


class ResidentSequence:
    """Same arithmetic, map held as per-voxel (sum, n) at a fixed origin; centroids are rounded to
    float32 as the GPU stores them, so this is the exact host twin of ``b2_slam_scan``."""

    def __init__(self, map_voxel, map_origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30,
                 relative_fitness=1e-6, relative_rmse=1e-6):
        self.map_voxel = map_voxel
        self.origin = np.asarray(map_origin, dtype=np.float64)
        self.ds_voxel = ds_voxel
        self.max_corr = max_corr
        self.max_iteration = max_iteration
        self.relative_fitness = relative_fitness
        self.relative_rmse = relative_rmse
        self.acc = {}  # (ix,iy,iz) -> [sx, sy, sz, n]
        self.pose = np.identity(4)
        self.last_result = None

    def map_points(self):
        keys = np.array(sorted(self.acc.keys()), dtype=np.int32).reshape(-1, 3)
        cent = np.array([np.array(self.acc[tuple(k)][:3]) / self.acc[tuple(k)][3] for k in keys]).reshape(-1, 3)
        return cent.astype(np.float32), keys

    def fuse(self, points, T):
        moved = cpu.transform(np.asarray(points, dtype=np.float64), T)
        # per-scan sums in input order (A.1), then merged into the accumulators
        k_all = cpu.voxel_keys(moved, self.map_voxel, origin=self.origin)
        order = {}
        for i, k in enumerate(map(tuple, k_all)):
            a = order.setdefault(k, [0.0, 0.0, 0.0, 0])
            a[0] += moved[i, 0]; a[1] += moved[i, 1]; a[2] += moved[i, 2]; a[3] += 1
        for k, a in order.items():
            m = self.acc.get(k)
            if m is None:
                self.acc[k] = list(a)
            else:
                m[0] += a[0]; m[1] += a[1]; m[2] += a[2]; m[3] += a[3]

    def process(self, points):
        pts32 = np.asarray(points, dtype=np.float32)
        if not self.acc:
            self.fuse(pts32, self.pose)
            return None
        src = pts32.astype(np.float64)
        if self.ds_voxel > 0:
            src, _, _ = cpu.voxel_downsample(src, self.ds_voxel)
            src = src.astype(np.float32).astype(np.float64)
        tgt, _ = self.map_points()
        res = cpu.registration_icp(src, tgt.astype(np.float64), self.max_corr, init=self.pose, mode="p2p",
                                   max_iteration=self.max_iteration, relative_fitness=self.relative_fitness,
                                   relative_rmse=self.relative_rmse)
        self.last_result = res
        self.pose = res.transformation
        self.fuse(pts32, self.pose)
        return res
