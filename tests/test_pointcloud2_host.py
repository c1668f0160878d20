"""Host side of SURVEY.md §8 row f3 (no GPU): the PointCloud2 message shape, the CDR / rosbridge encodings,
the rosbag2-shaped storage and the oracle restatement of sensor_msgs_py.point_cloud2 — known-answer vectors
written by hand from the published layouts, plus round trips."""
import base64
import struct

import numpy as np
import pytest

from lidar_slam_b200.plugin import bag
from lidar_slam_b200.plugin import pointcloud2 as pc2
from oracle import pointcloud2 as opc2


def _cloud(records: np.ndarray, dense=False) -> pc2.PointCloud2:
    c = pc2.create_cloud_xyz32(pc2.Header(pc2.Time(12, 34), "camera_link"), records)
    c.is_dense = dense
    return c


def test_create_cloud_xyz32_known_answer():
    pts = np.array([[1.0, 2.0, 3.0], [-0.5, 0.25, 1e-3]], dtype=np.float64)  # the reference passes float64 maps
    c = pc2.create_cloud_xyz32(pc2.Header(frame_id="map"), pts)
    assert (c.height, c.width, c.point_step, c.row_step, c.is_dense, c.is_bigendian) == (1, 2, 12, 24, False, False)
    assert [(f.name, f.offset, f.datatype, f.count) for f in c.fields] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1)]
    assert c.data == struct.pack("<6f", 1.0, 2.0, 3.0, -0.5, 0.25, 1e-3)
    assert c.data == opc2.create_cloud_xyz32_bytes(pts)


def test_oracle_read_points_fields_offsets_nan_and_dense_flag():
    # 20-byte records: [pad u32][x f32][intensity f32][y f32][z f32], one NaN record
    rows = [(7, 1.0, 9.0, 2.0, 3.0), (8, np.nan, 9.0, 5.0, 6.0), (9, 7.0, np.nan, 8.0, 9.0)]
    data = b"".join(struct.pack("<Iffff", *r) for r in rows)
    fields = [pc2.PointField("x", 4, 7, 1), pc2.PointField("intensity", 8, 7, 1), pc2.PointField("y", 12, 7, 1),
              pc2.PointField("z", 16, 7, 1)]
    c = pc2.PointCloud2(pc2.Header(), 1, 3, fields, False, 20, 60, data, False)
    got = opc2.read_points(c, ("x", "y", "z"), skip_nans=True)
    np.testing.assert_array_equal(opc2.xyz_array(got), np.array([[1, 2, 3], [7, 8, 9]], np.float32))  # NaN only in kept fields counts
    c.is_dense = True  # sensor_msgs_py skips the NaN scan for clouds flagged dense
    assert len(opc2.read_points(c, ("x", "y", "z"), skip_nans=True)) == 3
    be = pc2.PointCloud2(pc2.Header(), 1, 1, pc2._XYZ_FIELDS, True, 12, 12, struct.pack(">3f", 1.5, -2.0, 4.0), False)
    np.testing.assert_array_equal(opc2.xyz_array(opc2.read_points(be, ("x", "y", "z"))), [[1.5, -2.0, 4.0]])


def test_cdr_known_answer_and_round_trip():
    c = pc2.PointCloud2(pc2.Header(pc2.Time(1, 2), "map"), 1, 1, [pc2.PointField("x", 0, 7, 1)], False, 4, 4,
                        struct.pack("<f", 1.0), True)
    want = (b"\x00\x01\x00\x00"                      # encapsulation: CDR_LE
            + struct.pack("<iI", 1, 2)               # stamp
            + struct.pack("<I", 4) + b"map\x00"      # frame_id (length counts the NUL)
            + struct.pack("<II", 1, 1)               # height, width
            + struct.pack("<I", 1)                   # one field
            + struct.pack("<I", 2) + b"x\x00" + b"\x00\x00"  # name, padded to the next uint32
            + struct.pack("<I", 0)                   # offset
            + b"\x07" + b"\x00\x00\x00"              # datatype, padding
            + struct.pack("<I", 1)                   # count
            + b"\x00" + b"\x00\x00\x00"              # is_bigendian, padding
            + struct.pack("<II", 4, 4)               # point_step, row_step
            + struct.pack("<I", 4) + struct.pack("<f", 1.0)  # data
            + b"\x01")                               # is_dense
    assert pc2.serialize_cdr(c) == want
    assert pc2.deserialize_cdr(want) == c
    rng = np.random.default_rng(0)
    big = _cloud(rng.normal(size=(1000, 3)).astype(np.float32))
    assert pc2.deserialize_cdr(pc2.serialize_cdr(big)) == big
    with pytest.raises(ValueError):
        pc2.deserialize_cdr(b"\x05\x05\x00\x00")


def test_rosbridge_json_as_sent_by_the_phone():
    rec = np.array([[10, 20, 1.5], [11, 20, 1.25]], np.float32)
    msg = {"header": {"stamp": {"secs": 5, "nsecs": 6}, "frame_id": "camera_link"}, "height": 1, "width": 2,
           "fields": [{"name": "x", "offset": 0, "datatype": 7, "count": 1},
                      {"name": "y", "offset": 4, "datatype": 7, "count": 1},
                      {"name": "z", "offset": 8, "datatype": 7, "count": 1}],
           "is_bigendian": False, "point_step": 12, "row_step": 24,
           "data": base64.b64encode(rec.tobytes()).decode(), "is_dense": True}
    c = pc2.from_rosbridge(msg)  # ARDepthViewController.swift:136-150
    assert (c.header.stamp.sec, c.header.stamp.nanosec, c.width, c.point_step, c.is_dense) == (5, 6, 2, 12, True)
    np.testing.assert_array_equal(opc2.xyz_array(opc2.read_points(c, ("x", "y", "z"), True)), rec)
    assert pc2.from_rosbridge(pc2.to_rosbridge(c)) == c


def test_bag_round_trip_and_order(tmp_path):
    rng = np.random.default_rng(1)
    clouds = [_cloud(rng.normal(size=(n, 3)).astype(np.float32)) for n in (5, 0, 17)]
    uri = str(tmp_path / "input_bag")
    with bag.BagWriter(uri) as w:
        w.create_topic("/input_pointcloud")
        for k in (2, 0, 1):  # written out of order: the reader returns timestamp order
            w.write("/input_pointcloud", pc2.serialize_cdr(clouds[k]), 1000 + k)
        w.write_cloud("/global_map", clouds[2], 5000)
    r = bag.BagReader(uri)
    assert r.topics() == {"/input_pointcloud": (bag.POINTCLOUD2_TYPE, "cdr"), "/global_map": (bag.POINTCLOUD2_TYPE, "cdr")}
    got = list(r.clouds("/input_pointcloud"))
    assert [t for t, _ in got] == [1000, 1001, 1002]
    assert [c for _, c in got] == clouds
    assert len(list(r.messages())) == 4
    meta = (tmp_path / "input_bag" / "metadata.yaml").read_text()
    assert "storage_identifier: sqlite3" in meta and "message_count: 4" in meta
    with pytest.raises(FileExistsError):
        bag.BagWriter(uri)
    with pytest.raises(KeyError):
        w2 = bag.BagWriter(str(tmp_path / "other"))
        w2.write("/nope", b"", 0)


def test_cdr_and_rosbridge_round_trip_fuzz():
    """Property: any well-formed PointCloud2 survives CDR and rosbridge-JSON encoding (random field tables,
    names of every length mod 4 to exercise the alignment padding, payloads of any size)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    field = st.builds(pc2.PointField, st.text(alphabet="abcxyz_0123456789", min_size=0, max_size=9),
                      st.integers(0, 2 ** 31 - 1), st.integers(1, 8), st.integers(1, 16))
    cloud = st.builds(pc2.PointCloud2,
                      st.builds(pc2.Header, st.builds(pc2.Time, st.integers(-2 ** 31, 2 ** 31 - 1), st.integers(0, 2 ** 32 - 1)),
                                st.text(alphabet="abcdefghijklmnop/_", max_size=11)),
                      st.integers(0, 4), st.integers(0, 50), st.lists(field, max_size=5), st.booleans(),
                      st.integers(0, 64), st.integers(0, 4096), st.binary(max_size=300), st.booleans())

    @settings(max_examples=150, deadline=None)
    @given(cloud)
    def check(c):
        assert pc2.deserialize_cdr(pc2.serialize_cdr(c)) == c
        assert pc2.from_rosbridge(pc2.to_rosbridge(c)) == c

    check()
