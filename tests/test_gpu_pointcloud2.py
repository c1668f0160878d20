"""SURVEY.md §8 row f3 on the GPU: b2_ingest_pointcloud2 against the oracle restatement of
sensor_msgs_py.point_cloud2.read_points (bit-exact records, order preserved), and the ROS-free replay of a
rosbag2-shaped recording through the plugin."""
import multiprocessing
import os
import struct

import numpy as np
import pytest
import torch

from lidar_slam_b200.plugin import bag
from lidar_slam_b200.plugin import pointcloud2 as pc2
from oracle import pointcloud2 as opc2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, "3d-lidar-slam_b200", "config", "lidar_config.yaml")


def xyz(t):
    return t[:, :3].cpu().numpy()


def _records(n, seed, nan_every=0):
    rng = np.random.default_rng(seed)
    rec = np.stack([rng.integers(0, 256, n), rng.integers(0, 192, n), rng.uniform(0.3, 5.0, n)], axis=1).astype(np.float32)
    if nan_every:
        rec[::nan_every, seed % 3] = np.nan
    return rec


def test_ingest_xyz32_matches_read_points(engine):
    rec = _records(49152, 0, nan_every=7)
    c = pc2.create_cloud_xyz32(pc2.Header(), rec)
    want = opc2.xyz_array(opc2.read_points(c, ("x", "y", "z"), skip_nans=True))
    got = pc2.read_points_device(engine, c)
    assert 0 < len(want) < len(rec)
    np.testing.assert_array_equal(xyz(got), want)
    # flagged dense: sensor_msgs_py does not look for NaN, neither do we
    c.is_dense = True
    got = pc2.read_points_device(engine, c)
    np.testing.assert_array_equal(xyz(got), rec)
    # payload already on the device
    c.is_dense = False
    dev = torch.frombuffer(bytearray(c.data), dtype=torch.uint8).cuda()
    got = engine.ingest_pointcloud2(dev, len(rec), 12, (0, 4, 8))
    np.testing.assert_array_equal(xyz(got), want)


@pytest.mark.parametrize("layout", ["padded", "unaligned", "bigendian", "zyx"])
def test_ingest_general_layouts(engine, layout):
    rec = _records(5000, 3, nan_every=11)
    if layout == "padded":      # 32-byte records, extra fields around x/y/z
        step, offs, fmt = 32, (4, 12, 20), lambda r: struct.pack("<IfIfIfII", 1, r[0], 2, r[1], 3, r[2], 4, 5)
        big = False
    elif layout == "unaligned":  # 13-byte records: a leading byte shifts every field off 4-byte alignment
        step, offs, fmt = 13, (1, 5, 9), lambda r: b"\x7f" + struct.pack("<3f", *r)
        big = False
    elif layout == "bigendian":
        step, offs, fmt = 12, (0, 4, 8), lambda r: struct.pack(">3f", *r)
        big = True
    else:                        # fields stored z, y, x
        step, offs, fmt = 12, (8, 4, 0), lambda r: struct.pack("<3f", r[2], r[1], r[0])
        big = False
    data = b"".join(fmt(r) for r in rec)
    fields = [pc2.PointField(n, o, 7, 1) for n, o in zip("xyz", offs)]
    c = pc2.PointCloud2(pc2.Header(), 1, len(rec), fields, big, step, step * len(rec), data, False)
    want = opc2.xyz_array(opc2.read_points(c, ("x", "y", "z"), skip_nans=True))
    np.testing.assert_array_equal(xyz(pc2.read_points_device(engine, c)), want)


def test_ingest_fused_projection_and_edge_cases(engine, b2):
    from lidar_slam_b200 import synth

    rec = _records(20000, 5, nan_every=13)
    c = pc2.create_cloud_xyz32(pc2.Header(), rec)
    kept = opc2.xyz_array(opc2.read_points(c, ("x", "y", "z"), skip_nans=True))
    k = synth.intrinsics()
    for numpy2 in (False, True):
        got = pc2.read_points_device(engine, c, intrinsics=k, numpy2=numpy2)
        np.testing.assert_array_equal(xyz(got), opc2.project(kept, *k, numpy2=numpy2))
    empty = pc2.create_cloud_xyz32(pc2.Header(), np.zeros((0, 3), np.float32))
    assert pc2.read_points_device(engine, empty).shape == (0, 4)
    allnan = pc2.create_cloud_xyz32(pc2.Header(), np.full((100, 3), np.nan, np.float32))
    assert pc2.read_points_device(engine, allnan).shape == (0, 4)
    with pytest.raises(ValueError, match="no field"):
        pc2.read_points_device(engine, c, field_names=("x", "y", "intensity"))
    with pytest.raises(b2.B2Error, match="point_step"):
        engine.ingest_pointcloud2(c.data, 10, 12, (0, 4, 10))
    with pytest.raises(ValueError, match="shorter"):
        engine.ingest_pointcloud2(c.data[:100], 10, 12, (0, 4, 8))


def test_export_round_trip_through_pointcloud2(engine):
    """map -> create_cloud_xyz32 (device unpack) -> CDR -> read back on the device: identical points."""
    pts = np.random.default_rng(2).normal(size=(30000, 3)).astype(np.float32)
    p4 = engine.pack(pts)
    cloud = pc2.create_cloud_xyz32(pc2.Header(frame_id="map"), p4, engine=engine)
    assert cloud.data == opc2.create_cloud_xyz32_bytes(pts)
    back = pc2.read_points_device(engine, pc2.deserialize_cdr(pc2.serialize_cdr(cloud)))
    np.testing.assert_array_equal(xyz(back), pts)


def test_bag_replay_equals_queue_path(tmp_path):
    """A recorded /input_pointcloud bag replayed without ROS gives the state the live queue path gives."""
    from lidar_slam_b200 import synth
    from lidar_slam_b200.plugin import DataTransfer, set_algorithm

    scene = synth.make_corridor_scene(6, length=10.0)
    poses = synth.corridor_trajectory(scene, 4, seed=6, step=0.02)
    recs = [synth.scan_records(synth.render_depth(scene, p, seed=90 + k), keep_prob=0.6, seed=k).numpy()
            for k, p in enumerate(poses)]
    uri = str(tmp_path / "input_bags")
    with bag.BagWriter(uri) as w:
        for k, r in enumerate(recs):
            cloud = pc2.create_cloud_xyz32(pc2.Header(pc2.Time(k, 0), "camera_link"), r)
            cloud.is_dense = True  # as the phone flags it (ARDepthViewController.swift:148)
            w.write_cloud("/input_pointcloud", cloud, 10 ** 9 * k)

    def proc():
        dt = DataTransfer()
        return set_algorithm("b200_icp", CONFIG, dt, multiprocessing.Event(), mode="resident", max_iteration=30,
                             map_voxel=0.05), dt

    a, _ = proc()
    assert bag.replay_into(a, uri) == 4
    b, dtb = proc()
    for r in recs:
        dtb.pixel_depth_map_queue.put(r)
        b.process()
    assert a.point_clouds_in_map == b.point_clouds_in_map == 5
    np.testing.assert_array_equal(a.previous_transformation[-1], b.previous_transformation[-1])
    np.testing.assert_array_equal(a.global_map, b.global_map)
    # and the map leaves the way output_handler.py:71 sends it
    out = pc2.create_cloud_xyz32(pc2.Header(frame_id="map"), a.global_map)
    assert out.width == len(a.global_map) and out.point_step == 12
