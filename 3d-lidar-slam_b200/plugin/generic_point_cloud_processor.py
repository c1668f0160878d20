"""The processor plugin contract of the reference, restated.

Mirror of ``ProcessPointClouds`` in
src/slam/scripts/point_cloud_processors/generic_point_cloud_processor.py:11-159 —
same constructor signature, attribute names, ``process`` / ``reset`` behaviour and
abstract methods — so a processor written against this class drops into the
reference unchanged (INTEGRATION.md shows the two-line registration).  When the
reference package is importable its own base class is used instead, which makes
``B200ICPProcessor`` a genuine subclass of the reference's ABC.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod

import numpy as np
import yaml

try:  # inside the reference's ROS 2 workspace: subclass the real thing
    from scripts.point_cloud_processors.generic_point_cloud_processor import (  # type: ignore
        ProcessPointClouds as _ReferenceBase,
    )
except Exception:  # standalone (this repository, tests, bench)
    _ReferenceBase = None


class _ProcessPointCloudsMirror(ABC):
    """Per-scan driver + camera model shared by every registration algorithm."""

    def __init__(self, config_path: str, reset_event, data_transfer, logger=None):
        self.logger = logger
        self.WIDTH = self.HEIGHT = None
        self.ref_width = self.ref_height = None
        self.fx = self.fy = self.cx = self.cy = None
        self.load_config(config_path)

        self.global_map = None
        self.point_clouds_in_map = 0
        self.data_transfer = data_transfer
        # generic_point_cloud_processor.py:50-52 reads a map left in the output queue; the
        # attribute it copies from does not exist on DataTransfer, so only the probe is kept.
        with self.data_transfer.global_map_lock:
            self.data_transfer.global_map_queue.empty()
        self.previous_transformation = [np.identity(4)]
        self.start_time = None
        self.reset_event = reset_event
        self.logger.info(f"Using {self.__class__.__name__} for point cloud processing")

    def load_config(self, config_path: str) -> None:
        """Camera intrinsics from YAML, rescaled from the reference resolution to WIDTH x HEIGHT
        (generic_point_cloud_processor.py:62-90)."""
        with open(config_path, "r") as fh:
            cam = yaml.safe_load(fh)["camera"]
        self.WIDTH, self.HEIGHT = cam["width"], cam["height"]
        self.ref_width, self.ref_height = cam["ref_width"], cam["ref_height"]
        sx, sy = self.WIDTH / self.ref_width, self.HEIGHT / self.ref_height
        self.fx, self.fy = cam["fx"] * sx, cam["fy"] * sy
        self.cx, self.cy = cam["cx"] * sx, cam["cy"] * sy
        self.logger.info(
            f"Loaded camera parameters: WIDTH={self.WIDTH}, HEIGHT={self.HEIGHT}, fx={self.fx}, fy={self.fy}")

    def project_pixel_to_3d(self, points) -> np.ndarray:
        """(x_px, y_px, depth_m) records -> float32 [N,3] metres, pinhole model
        (generic_point_cloud_processor.py:93-108)."""
        out = np.zeros((len(points), 3), dtype=np.float32)
        for i, (x, y, z) in enumerate(points):
            out[i] = [(x - self.cx) * z / self.fx, (y - self.cy) * z / self.fy, z]
        return out

    def process(self) -> None:
        """Take one scan from the input queue, project it and hand it to the algorithm
        (generic_point_cloud_processor.py:111-126)."""
        scan = self.data_transfer.pixel_depth_map_queue.get()
        self.logger.debug(f"{len(scan)} points received from queue")
        if scan is not None:
            self.logger.info(f"pcs processed so far: {self.point_clouds_in_map}")
            pts = self.project_pixel_to_3d(scan)
            del scan
            self.construct_global_map(pts)
            self.point_clouds_in_map += 1

    def reset(self) -> None:
        """generic_point_cloud_processor.py:128-138."""
        self.logger.info("Resetting processor")
        self.global_map = None
        self.point_clouds_in_map = 0
        self.previous_transformation = [np.identity(4)]
        self.start_time = None
        self.reset_event.clear()
        time.sleep(1)

    @abstractmethod
    def construct_global_map(self, points: np.ndarray) -> None:
        """Register ``points`` (float32 [N,3], metres, sensor frame) and merge them into ``global_map``."""

    @abstractmethod
    def downsample_global_map(self) -> np.ndarray:
        """A thinned copy of the global map for display."""


ProcessPointClouds = _ReferenceBase if _ReferenceBase is not None else _ProcessPointCloudsMirror
