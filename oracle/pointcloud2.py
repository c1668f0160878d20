"""TEST INFRASTRUCTURE — CPU restatement of the two sensor_msgs_py.point_cloud2 helpers the reference calls
(PARITY UNPINNED: sensor_msgs_py is a ROS 2 Humble dependency that is absent from /root/reference and from this
image; the behaviour below restates its published implementation, anchored on the reference's call sites).

* ``read_points(cloud, field_names, skip_nans)`` — src/slam/slam/input_handler.py:65-66
  ``pc2.read_points(msg, field_names=("x", "y", "z"), skip_nans=True)``: view the payload as a structured
  array (one record per point_step bytes, a named scalar per field at its offset), keep the requested fields,
  byte-swap when the message's endianness differs from the host's, and — only when the cloud is NOT flagged
  ``is_dense`` — drop every record holding a NaN in one of the kept fields.
* ``create_cloud_xyz32(header, points)`` — output_handler.py:71, generic_handler.py:41: float32 x/y/z at
  offsets 0/4/8, point_step 12, height 1, width N, host byte order, ``is_dense`` False.

Only tests/ may import this module; the product path is b2_ingest_pointcloud2 / b2_unpack_xyz on the GPU.
"""
from __future__ import annotations

import sys

import numpy as np

_NP = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}


def dtype_from_fields(fields, point_step=None) -> np.dtype:
    names, formats, offsets = [], [], []
    for f in fields:
        base = np.dtype(_NP[f.datatype])
        for k in range(f.count):
            names.append(f.name if f.count == 1 else f"{f.name}_{k}")
            formats.append(base)
            offsets.append(f.offset + k * base.itemsize)
    spec = {"names": names, "formats": formats, "offsets": offsets}
    if point_step is not None:
        spec["itemsize"] = point_step
    return np.dtype(spec)


def read_points(cloud, field_names=None, skip_nans=False) -> np.ndarray:
    n = cloud.width * cloud.height
    points = np.ndarray(shape=(n,), dtype=dtype_from_fields(cloud.fields, cloud.point_step), buffer=bytearray(cloud.data))
    if field_names is not None:
        points = points[list(field_names)]
    if (sys.byteorder != "little") != bool(cloud.is_bigendian):
        points = points.byteswap()
    if skip_nans and not cloud.is_dense:
        keep = np.ones(len(points), dtype=bool)
        for name in points.dtype.names:
            keep &= ~np.isnan(points[name])
        points = points[keep]
    return points


def xyz_array(points) -> np.ndarray:
    """structured (x, y, z) -> float32 [n,3]."""
    return np.stack([np.asarray(points[n], dtype=np.float32) for n in points.dtype.names[:3]], axis=1).reshape(-1, 3)


def create_cloud_xyz32_bytes(points) -> bytes:
    """The ``data`` payload create_cloud_xyz32 produces for an [N,3] array."""
    return np.ascontiguousarray(np.asarray(points).reshape(-1, 3), dtype=np.float32).tobytes()


def project(records: np.ndarray, fx, fy, cx, cy, numpy2: bool) -> np.ndarray:
    """generic_point_cloud_processor.py:93-108 on float32 [n,3] records (NumPy-2 float32 / NumPy-1 float64 rules)."""
    r = np.asarray(records, dtype=np.float32)
    if numpy2:
        x = (r[:, 0] - np.float32(cx)) * r[:, 2] / np.float32(fx)
        y = (r[:, 1] - np.float32(cy)) * r[:, 2] / np.float32(fy)
        return np.stack([x, y, r[:, 2]], axis=1).astype(np.float32)
    d = r.astype(np.float64)
    x = (d[:, 0] - cx) * d[:, 2] / fx
    y = (d[:, 1] - cy) * d[:, 2] / fy
    return np.stack([x, y, d[:, 2]], axis=1).astype(np.float32)
