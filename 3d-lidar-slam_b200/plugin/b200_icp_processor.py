"""B200ICPProcessor — drop-in replacement for the reference's ``ICPProcessor``
(src/slam/scripts/point_cloud_processors/icp_processor.py:18-160).

Same constructor, same ``construct_global_map`` / ``downsample_global_map`` /
``process`` / ``reset`` contract, same hard-coded defaults (voxel 0.02 m, max
correspondence 2.5 voxels, point-to-point, 500 iterations, 1e-6 / 1e-6), but every
Open3D call is replaced by a call into libb2slam.so (hand-written sm_100a kernels).
There is no CPU fallback: constructing the processor without a CUDA device raises.

Two map modes:

``mode="resident"`` (default, the fast path)
    The global map is a GPU-resident voxel hash (b2_map.cu).  Per scan one C-ABI
    call (``b2_slam_scan``) uploads the scan, voxel-down-samples it, runs ICP against
    the resident map seeded with the previous pose and fuses the raw scan — O(scan)
    work, nothing O(map), no host round trip inside the ICP loop.  The map is
    anchored at a fixed voxel origin (the reference re-anchors at the growing map's
    min bound every scan, icp_processor.py:91) and ``global_map`` is the voxel
    centroids at ``map_voxel`` resolution, materialised on the host lazily.
    ``construct_global_map`` / ``process`` only ENQUEUE a scan (the pose prior of the
    next registration lives on the device, so nothing waits for the host); the poses
    are read back when somebody looks: ``previous_transformation`` (any access),
    ``last_result``, ``global_map`` — or after ``MAX_PENDING`` scans.  The reference
    reads ``previous_transformation`` only inside ``align_point_clouds_with_icp``
    (icp_processor.py:106,118), so the deferral is invisible to its callers; an error
    of a deferred scan (map full, non-finite points) is raised at that read instead
    of at the call.  A pageable scan array has been copied when the call returns; a
    pinned (cudaHostAlloc / registered) array is read in place and must stay
    unchanged until the poses have been read.

``mode="reference"``
    The reference's exact sequence (icp_processor.py:43-75,77-119) on GPU
    primitives: the raw map is kept point-for-point (on the device), both clouds
    are re-voxelised per scan with Open3D's per-cloud origin, the raw scan is
    transformed and PREPENDED.  Used by the parity tests against the CPU oracle.
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from .generic_point_cloud_processor import ProcessPointClouds
from .logging_shim import get_logger


class _PoseLog(list):
    """``previous_transformation`` of a resident-mode processor: a list that collects the poses of the scans
    still in flight on the device before it lets anyone look (icp_processor.py:106,118 only ever reads [-1] and
    appends; the GPU keeps its own copy of the pose prior, so registration never waits for the host)."""

    def __init__(self, items, owner):
        super().__init__(items)
        self._owner = owner

    def _sync(self):
        owner = self._owner
        if owner is not None and owner._pending_first:
            owner.flush_results()

    def __getitem__(self, i):
        self._sync()
        return super().__getitem__(i)

    def __len__(self):
        self._sync()
        return super().__len__()

    def __iter__(self):
        self._sync()
        return super().__iter__()

    def __repr__(self):
        self._sync()
        return super().__repr__()

    def __eq__(self, other):
        self._sync()
        return list.__eq__(self, other)

    __hash__ = None


class B200ICPProcessor(ProcessPointClouds):
    # resident mode submits scans asynchronously; results are read back when somebody looks at them
    # (previous_transformation, last_result, global_map) or when this many scans are pending
    MAX_PENDING = 256

    def __init__(self, config_path: str, reset_event, data_transfer, *, mode: str = "resident",
                 voxel_size: float = 0.02, max_correspondence_distance: float | None = None,
                 max_iteration: int = 500, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
                 map_voxel: float | None = None, map_origin=(0.0, 0.0, 0.0), map_capacity_cells: int = 1 << 22,
                 maintenance: str = "full", numpy2_projection: bool | None = None, device: int = 0,
                 logger=None):
        if mode not in ("resident", "reference"):
            raise ValueError(f"unknown mode {mode!r}")
        if maintenance not in ("voxel", "off", "full"):
            raise ValueError(f"unknown maintenance {maintenance!r}")
        self._map_cache = None
        self._pending_first = []   # one flag per deferred scan: was it the first scan of the map?
        self._mode = mode
        self._engine = _lib.Engine(device)  # raises without CUDA / without the built library
        self._voxel = float(voxel_size)
        self._max_corr = float(max_correspondence_distance) if max_correspondence_distance else 2.5 * self._voxel
        self._params = _lib.make_params(self._max_corr, max_iteration, _lib.P2P, relative_fitness, relative_rmse)
        self._map_voxel = float(map_voxel) if map_voxel else self._voxel
        self._map_origin = np.asarray(map_origin, dtype=np.float64)
        self._map_capacity = int(map_capacity_cells)
        self._maintenance = maintenance
        self._numpy2 = (int(np.__version__.split(".")[0]) >= 2) if numpy2_projection is None else bool(numpy2_projection)
        self._raw_map = None       # reference mode: CUDA float32 [N,4], newest scan first
        self._resident_ready = False
        self._last_result = None   # b2_icp_result of the most recent registration
        self._poses = _PoseLog([np.identity(4)], self)
        super().__init__(config_path, reset_event, data_transfer,
                         logger if logger is not None else get_logger("b200 icp processor"))

    # -- results that may still be in flight -----------------------------------
    @property
    def previous_transformation(self):
        return self._poses

    @previous_transformation.setter
    def previous_transformation(self, value):  # (the base class assigns [identity] in __init__ and reset())
        self._pending_first = []
        self._poses = _PoseLog(list(value), self)

    @property
    def last_result(self):
        if self._pending_first:
            self.flush_results()
        return self._last_result

    @last_result.setter
    def last_result(self, value):
        self._last_result = value

    # -- resuming from a resident map ------------------------------------------
    @property
    def engine(self):
        """The ``_lib.Engine`` (C-ABI handle + stream) this processor drives."""
        return self._engine

    def adopt_resident_map(self) -> None:
        """Continue from a map that is already resident in this processor's engine — filled through
        ``engine.map_init`` + ``engine.map_fuse`` / ``map_merge_records``, e.g. the map of an earlier session or
        the merged partial maps of a sharded replay.  The next scan is registered against it instead of becoming
        the map ("first scan is the map", icp_processor.py:53-56, applies to an empty map only)."""
        if self._mode != "resident":
            raise ValueError("adopt_resident_map() needs mode='resident'")
        self._resident_ready = True
        self._map_cache = None
        if self.point_clouds_in_map == 0:
            self.point_clouds_in_map = 1

    def set_pose_prior(self, T) -> None:
        """Seed the next registration (the role of previous_transformation[-1], icp_processor.py:106)."""
        T = np.array(T, dtype=np.float64).reshape(4, 4)
        self.flush_results()
        self._engine.set_pose(T)
        list.append(self._poses, T)

    # -- global_map: lazy device -> host materialisation ----------------------
    @property
    def global_map(self):
        if self._map_cache is None:
            if self._mode == "resident" and self._resident_ready and self.point_clouds_in_map > 0:
                self.flush_results()
                pts, _ = self._engine.map_export(want_keys=False)
                self._map_cache = self._engine.unpack(pts) if pts.shape[0] else np.zeros((0, 3), np.float32)
            elif self._mode == "reference" and self._raw_map is not None:
                self._map_cache = self._engine.unpack(self._raw_map)
        return self._map_cache

    @global_map.setter
    def global_map(self, value):
        self._map_cache = value

    # -- per-scan entry points -----------------------------------------------
    def project_pixel_to_3d(self, points) -> np.ndarray:
        """Vectorised statement of generic_point_cloud_processor.py:93-108 (float32 [N,3])."""
        rec = _records_to_array(points)
        if self._numpy2:
            r = rec.astype(np.float32)
            x = (r[:, 0] - np.float32(self.cx)) * r[:, 2] / np.float32(self.fx)
            y = (r[:, 1] - np.float32(self.cy)) * r[:, 2] / np.float32(self.fy)
            return np.stack([x, y, r[:, 2]], axis=1)
        r = rec.astype(np.float64)
        x = (r[:, 0] - self.cx) * r[:, 2] / self.fx
        y = (r[:, 1] - self.cy) * r[:, 2] / self.fy
        return np.stack([x, y, r[:, 2]], axis=1).astype(np.float32)

    def process(self) -> None:
        """ProcessPointClouds.process (generic_point_cloud_processor.py:111-126) with the projection fused
        into the upload: the (x_px, y_px, depth) records go to the device as they are and the front half of the scan
        step back-projects them there (``b2_slam_scan_records``; ``b2_project_pixels`` +
        ``b2_slam_scan_dev`` in reference mode), so the 49k-iteration Python loop of the reference — and any host-side
        array of metric points — disappears from the per-scan path."""
        self.process_records(self.data_transfer.pixel_depth_map_queue.get())

    def process_records(self, scan, defer_result: bool = False) -> None:
        """The body of ``process()`` for a scan that has already been dequeued (``ProcessLoop`` drains the
        input queue in batches).  In resident mode a scan is only enqueued, whatever
        ``defer_result`` says (kept for callers written against round 1); ``flush_results()`` appends the poses
        of all pending scans to ``previous_transformation`` with one device synchronisation for the batch."""
        del defer_result
        self._process_records(scan)

    def flush_results(self) -> int:
        """Collect the results of the scans submitted with ``defer_result=True``; returns how many."""
        pending, self._pending_first = self._pending_first, []  # (cleared first: a failed read-back must not
        n = len(pending)                                         #  be retried against later scans' records)
        if n == 0:
            return 0
        results = self._engine.slam_results(n)
        for first, res in zip(pending, results):
            if first:
                continue  # the first scan is the map frame: no registration (icp_processor.py:53-56)
            self._last_result = res
            list.append(self._poses, res.matrix())
        return n

    def _process_records(self, scan) -> None:
        if scan is not None:
            self.logger.debug(f"{len(scan)} points received from queue")
            self.logger.info(f"pcs processed so far: {self.point_clouds_in_map}")
            rec = np.ascontiguousarray(_records_to_array(scan), dtype=np.float32)
            del scan
            if self._mode == "resident":
                # one call: upload, projection, down-sample (front stream) -> ICP -> fuse (main stream)
                self._map_cache = None
                self._construct_resident(None, records=rec)
                self.point_clouds_in_map += 1
                return
            pts4 = self._engine.project_pixels(rec, self.fx, self.fy, self.cx, self.cy, numpy2=self._numpy2)
            self._register_device_scan(pts4)

    def process_cloud(self, cloud) -> None:
        """One scan given as the PointCloud2 it arrived in (``plugin.pointcloud2.PointCloud2``; the wire format
        of input_handler.py:65-66): the payload bytes go to the device, where ``b2_ingest_pointcloud2`` does
        what ``read_points(msg, ("x","y","z"), skip_nans=True)`` + ``project_pixel_to_3d`` do on the host in
        the reference.  Same state changes as ``process()``; used for ROS-free replay (``plugin.bag``)."""
        from . import pointcloud2 as pc2

        self.logger.info(f"pcs processed so far: {self.point_clouds_in_map}")
        pts4 = pc2.read_points_device(self._engine, cloud, ("x", "y", "z"), skip_nans=True,
                                      intrinsics=(self.fx, self.fy, self.cx, self.cy), numpy2=self._numpy2)
        self._register_device_scan(pts4)

    def _register_device_scan(self, pts4) -> None:
        self._map_cache = None
        if self._mode == "resident":
            self._construct_resident(None, pts4)
        else:
            self._construct_reference(pts4)
        self.point_clouds_in_map += 1

    def construct_global_map(self, points: np.ndarray) -> None:
        """ICPProcessor.construct_global_map (icp_processor.py:43-75)."""
        pts = np.ascontiguousarray(points, dtype=np.float32)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError(f"expected [N,3] points, got {pts.shape}")
        self._map_cache = None
        if self._mode == "resident":
            self._construct_resident(pts)
        else:
            self._construct_reference(self._engine.pack(pts))

    def downsample_global_map(self) -> np.ndarray:
        """ICPProcessor.downsample_global_map (icp_processor.py:150-160) returns the map as is."""
        return self.global_map

    def reset(self) -> None:
        self._pending_first = []
        if self._resident_ready:
            self._engine.synchronize()
            self._engine.map_reset()
        self._raw_map = None
        self._map_cache = None
        self._last_result = None
        super().reset()

    # -- resident mode --------------------------------------------------------
    def _construct_resident(self, pts, pts4=None, records=None) -> None:
        eng = self._engine
        if not self._resident_ready:
            eng.map_init(self._map_voxel, self._map_origin, self._max_corr, self._map_capacity)
            self._resident_ready = True
        first = self.point_clouds_in_map == 0
        # enqueue only: upload, down-sample, ICP seeded with the pose kept on the device, fuse.  The pose comes
        # back when somebody reads previous_transformation / last_result / global_map (or MAX_PENDING is reached);
        # a pageable `pts` has been copied into the library's pinned ring when the call returns
        if records is not None:
            eng.slam_scan_records(records, self.fx, self.fy, self.cx, self.cy, self._numpy2, self._voxel, self._params,
                                  want_result=False)
        elif pts4 is not None:
            eng.slam_scan_dev(pts4, self._voxel, self._params, want_result=False)
        else:
            eng.slam_scan(pts, self._voxel, self._params, want_result=False)
        self._pending_first.append(first)
        if first:
            self.point_clouds_in_map += 1  # icp_processor.py:54 (the caller increments once more)
        if len(self._pending_first) >= self.MAX_PENDING:
            self.flush_results()

    # -- reference mode -------------------------------------------------------
    def _construct_reference(self, scan4) -> None:
        eng = self._engine
        torch = eng.torch
        if self.point_clouds_in_map == 0:
            self.point_clouds_in_map += 1
            self._raw_map = scan4
            return
        source_down, _ = eng.voxel_downsample(scan4, self._voxel, want_keys=False)
        target_down, _ = eng.voxel_downsample(self._raw_map, self._voxel, want_keys=False)
        # icp_processor.py:94-99 estimates normals on both clouds here and never uses them
        # (point-to-point estimation): skipped, the result does not depend on them.
        res, _ = eng.icp(source_down, target_down, self._params, init=self.previous_transformation[-1])
        self._last_result = res
        T = res.matrix()
        list.append(self._poses, T)
        moved = eng.transform(scan4, T)
        with torch.cuda.stream(eng.stream):
            self._raw_map = torch.cat([moved, self._raw_map], dim=0)
        self._do_stuff()

    def _do_stuff(self) -> None:
        """ICPProcessor.do_stuff (icp_processor.py:121-148): every 10th scan (or whenever the input queue is
        empty) re-voxelise the map at 1 mm and drop statistical outliers (45 neighbours, 3.5 sigma); every
        40th scan use (80, 2.0) followed by the radius filter (more than 8 points within 2.3 cm)."""
        if self._maintenance == "off" or self.point_clouds_in_map < 10:
            return
        if self.point_clouds_in_map % 10 == 0 or self.data_transfer.pixel_depth_map_queue.empty():
            eng = self._engine
            m, _ = eng.voxel_downsample(self._raw_map, 0.001, want_keys=False)
            m = m.contiguous()
            if self._maintenance == "full":
                if self.point_clouds_in_map % 40 == 0:
                    keep, _ = eng.statistical_outlier_mask(m, 80, 2.0)
                    m = eng.select_points(m, keep).contiguous()
                    keep, _ = eng.radius_outlier_mask(m, 8, 0.023)
                    m = eng.select_points(m, keep).contiguous()
                else:
                    keep, _ = eng.statistical_outlier_mask(m, 45, 3.5)
                    m = eng.select_points(m, keep).contiguous()
            self._raw_map = m


def _records_to_array(points) -> np.ndarray:
    a = np.asarray(points)
    if a.dtype.names:  # structured array from sensor_msgs_py.point_cloud2.read_points (input_handler.py:65)
        a = np.stack([a[n] for n in a.dtype.names[:3]], axis=1)
    return np.asarray(a).reshape(-1, 3)
