"""PointCloud2-shaped ingest and export without ROS.

The reference moves every scan and every map as ``sensor_msgs/msg/PointCloud2``: the phone sends
12-byte ``(x_px, y_px, depth)`` float32 records through rosbridge
(ios/.../ARDepthViewController.swift:136-150), the subscriber turns the payload into a structured
array with ``sensor_msgs_py.point_cloud2.read_points(msg, field_names=("x","y","z"), skip_nans=True)``
(src/slam/slam/input_handler.py:65-66), and maps leave through
``point_cloud2.create_cloud_xyz32(header, points)`` (output_handler.py:57-74, generic_handler.py:33-44).

This module carries the same message shape as plain dataclasses, so offline replay can feed the GPU
path on a box without ROS:

* ``read_points_device``  — the payload goes to the GPU as bytes and ``b2_ingest_pointcloud2`` decodes
  the three fields, drops NaN records (order preserved) and, given the intrinsics, back-projects in
  the same pass; nothing is parsed on the host.
* ``create_cloud_xyz32``  — wraps an ``[N,3]`` array or a device ``[N,4]`` tensor (unpacked on the GPU by
  ``b2_unpack_xyz``) in a PointCloud2 with the reference's field layout.
* ``from_rosbridge`` / ``to_rosbridge`` — the JSON form rosbridge carries (base64 ``data``).
* ``serialize_cdr`` / ``deserialize_cdr`` — the CDR encoding rosbag2 stores (see ``bag.py``).

sensor_msgs_py / rosbag2 are not installed here: the formats are restated from their
published definitions and pinned only by this repo's own round-trip and known-answer tests.
"""
from __future__ import annotations

import base64
import struct
import sys
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np


@dataclass
class Time:
    sec: int = 0
    nanosec: int = 0


@dataclass
class Header:
    stamp: Time = field(default_factory=Time)
    frame_id: str = ""


@dataclass
class PointField:
    INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = range(1, 9)
    name: str = ""
    offset: int = 0
    datatype: int = 7
    count: int = 1


@dataclass
class PointCloud2:
    header: Header = field(default_factory=Header)
    height: int = 1
    width: int = 0
    fields: List[PointField] = field(default_factory=list)
    is_bigendian: bool = False
    point_step: int = 0
    row_step: int = 0
    data: bytes = b""
    is_dense: bool = False

    def __len__(self) -> int:
        return self.height * self.width


_XYZ_FIELDS = [PointField("x", 0, PointField.FLOAT32, 1), PointField("y", 4, PointField.FLOAT32, 1),
               PointField("z", 8, PointField.FLOAT32, 1)]


def _xyz_offsets(cloud: PointCloud2, field_names: Sequence[str]):
    offs = []
    for name in field_names:
        f = next((f for f in cloud.fields if f.name == name), None)
        if f is None:
            raise ValueError(f"PointCloud2 has no field {name!r}")
        if f.datatype != PointField.FLOAT32 or f.count != 1:
            raise ValueError(f"field {name!r} is not a scalar float32 (datatype {f.datatype}, count {f.count})")
        offs.append(f.offset)
    return offs


def read_points_device(engine, cloud: PointCloud2, field_names=("x", "y", "z"), skip_nans: bool = True,
                       intrinsics=None, numpy2: bool = False):
    """``read_points(cloud, field_names, skip_nans)`` evaluated on the GPU: returns a CUDA float32 ``[m,4]``
    tensor (xyz + pad) of the kept records, in message order.  ``intrinsics=(fx, fy, cx, cy)`` fuses the
    pixel->metric projection of ``ProcessPointClouds.project_pixel_to_3d`` into the same pass."""
    if len(field_names) != 3:
        raise ValueError("exactly three fields are read (x, y, z)")
    offs = _xyz_offsets(cloud, field_names)
    n = cloud.height * cloud.width
    if cloud.height > 1 and cloud.row_step != cloud.width * cloud.point_step:
        raise ValueError("organised clouds with row padding are not supported")
    # sensor_msgs_py only looks for NaN when the sender did not flag the cloud dense
    skip = bool(skip_nans) and not cloud.is_dense
    return engine.ingest_pointcloud2(cloud.data, n, cloud.point_step, offs, cloud.is_bigendian, skip, intrinsics,
                                     numpy2)


def create_cloud_xyz32(header: Header, points, engine=None) -> PointCloud2:
    """``point_cloud2.create_cloud_xyz32``: height 1, width N, float32 x/y/z at offsets 0/4/8, point_step 12,
    host byte order, ``is_dense`` False.  ``points``: anything ``np.asarray`` turns into ``[N,3]`` (the
    reference passes float64 maps; they are cast to float32 as the ROS helper does), or a CUDA ``[N,4]``
    tensor, which is unpacked to 12-byte records on the device (needs ``engine``)."""
    if hasattr(points, "is_cuda") and points.is_cuda:
        if engine is None:
            raise ValueError("a device tensor needs the engine that owns it")
        arr = engine.unpack(points)
    else:
        arr = np.asarray(points)
        if arr.dtype.names:
            arr = np.stack([arr[n] for n in arr.dtype.names[:3]], axis=1)
        arr = np.ascontiguousarray(arr.reshape(-1, 3), dtype=np.float32)
    n = int(arr.shape[0])
    return PointCloud2(header=header, height=1, width=n, fields=[PointField(f.name, f.offset, f.datatype, f.count)
                                                                 for f in _XYZ_FIELDS],
                       is_bigendian=sys.byteorder != "little", point_step=12, row_step=12 * n,
                       data=arr.tobytes(), is_dense=False)


# ----------------------------------------------------------------------------- rosbridge JSON
def from_rosbridge(msg: dict) -> PointCloud2:
    """The dict rosbridge delivers for a PointCloud2 (``data`` base64-encoded; stamp keys ``secs``/``nsecs`` as
    the iOS client sends them, or ``sec``/``nanosec``)."""
    h = msg.get("header", {})
    st = h.get("stamp", {})
    stamp = Time(int(st.get("sec", st.get("secs", 0))), int(st.get("nanosec", st.get("nsecs", 0))))
    data = msg.get("data", b"")
    if isinstance(data, str):
        data = base64.b64decode(data)
    elif isinstance(data, list):
        data = bytes(data)
    fields = [PointField(f["name"], int(f["offset"]), int(f["datatype"]), int(f.get("count", 1)))
              for f in msg.get("fields", [])]
    return PointCloud2(Header(stamp, h.get("frame_id", "")), int(msg.get("height", 1)), int(msg.get("width", 0)),
                       fields, bool(msg.get("is_bigendian", False)), int(msg.get("point_step", 0)),
                       int(msg.get("row_step", 0)), bytes(data), bool(msg.get("is_dense", False)))


def to_rosbridge(cloud: PointCloud2) -> dict:
    return {"header": {"stamp": {"sec": cloud.header.stamp.sec, "nanosec": cloud.header.stamp.nanosec},
                       "frame_id": cloud.header.frame_id},
            "height": cloud.height, "width": cloud.width,
            "fields": [{"name": f.name, "offset": f.offset, "datatype": f.datatype, "count": f.count}
                       for f in cloud.fields],
            "is_bigendian": cloud.is_bigendian, "point_step": cloud.point_step, "row_step": cloud.row_step,
            "data": base64.b64encode(cloud.data).decode("ascii"), "is_dense": cloud.is_dense}


# ----------------------------------------------------------------------------- CDR (rosbag2 "cdr")
class _Writer:
    def __init__(self):
        self.b = bytearray(b"\x00\x01\x00\x00")  # encapsulation header: CDR little endian, no options

    def _align(self, n):
        pad = (-(len(self.b) - 4)) % n
        self.b += b"\x00" * pad

    def put(self, fmt, v, size):
        self._align(size)
        self.b += struct.pack("<" + fmt, v)

    def string(self, s):
        raw = s.encode("utf-8") + b"\x00"
        self.put("I", len(raw), 4)
        self.b += raw

    def octets(self, data):
        self.put("I", len(data), 4)
        self.b += data


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        if len(self.b) < 4 or self.b[0] != 0 or self.b[1] not in (0, 1):
            raise ValueError("not a plain CDR buffer")
        self.e = "<" if self.b[1] == 1 else ">"
        self.p = 4

    def _align(self, n):
        self.p += (-(self.p - 4)) % n

    def get(self, fmt, size):
        self._align(size)
        (v,) = struct.unpack_from(self.e + fmt, self.b, self.p)
        self.p += size
        return v

    def string(self):
        n = self.get("I", 4)
        s = bytes(self.b[self.p:self.p + n - 1]).decode("utf-8") if n else ""
        self.p += n
        return s

    def octets(self):
        n = self.get("I", 4)
        d = bytes(self.b[self.p:self.p + n])
        self.p += n
        return d


def serialize_cdr(cloud: PointCloud2) -> bytes:
    """sensor_msgs/msg/PointCloud2 in XCDR1 little endian, the ``serialization_format="cdr"`` rosbag2 stores
    (generic_handler.py:43 ``serialize_message(cloud_msg)``)."""
    w = _Writer()
    w.put("i", cloud.header.stamp.sec, 4)
    w.put("I", cloud.header.stamp.nanosec, 4)
    w.string(cloud.header.frame_id)
    w.put("I", cloud.height, 4)
    w.put("I", cloud.width, 4)
    w.put("I", len(cloud.fields), 4)
    for f in cloud.fields:
        w.string(f.name)
        w.put("I", f.offset, 4)
        w.put("B", f.datatype, 1)
        w.put("I", f.count, 4)
    w.put("B", 1 if cloud.is_bigendian else 0, 1)
    w.put("I", cloud.point_step, 4)
    w.put("I", cloud.row_step, 4)
    w.octets(cloud.data)
    w.put("B", 1 if cloud.is_dense else 0, 1)
    return bytes(w.b)


def deserialize_cdr(buf) -> PointCloud2:
    r = _Reader(buf)
    sec = r.get("i", 4)
    nsec = r.get("I", 4)
    frame = r.string()
    height = r.get("I", 4)
    width = r.get("I", 4)
    fields = []
    for _ in range(r.get("I", 4)):
        name = r.string()
        off = r.get("I", 4)
        dt = r.get("B", 1)
        cnt = r.get("I", 4)
        fields.append(PointField(name, off, dt, cnt))
    big = bool(r.get("B", 1))
    step = r.get("I", 4)
    row = r.get("I", 4)
    data = r.octets()
    dense = bool(r.get("B", 1))
    return PointCloud2(Header(Time(sec, nsec), frame), height, width, fields, big, step, row, data, dense)
