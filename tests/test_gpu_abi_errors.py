"""The C ABI's error behaviour through raw ctypes calls (include/b2slam.h: every function returns B2_OK or a
negative B2_ERR_* code and leaves a message in b2_last_error; nothing crashes on a bad argument)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

INVALID, STATE = -1, -5


@pytest.fixture()
def raw(b2):
    eng = b2.Engine(0)
    yield eng, eng.lib, eng.h
    eng.close()


def _err(lib, h):
    return lib.b2_last_error(h).decode()


def test_null_and_negative_arguments_are_rejected(raw, b2):
    eng, lib, h = raw
    pts = eng.pack(np.random.default_rng(0).normal(size=(100, 3)).astype(np.float32))
    out = eng.empty_points(100)
    n = C.c_int(0)
    assert lib.b2_voxel_downsample(h, None, 100, 0.02, None, None, out.data_ptr(), None, C.byref(n)) == INVALID
    assert lib.b2_voxel_downsample(h, pts.data_ptr(), -1, 0.02, None, None, out.data_ptr(), None, C.byref(n)) == INVALID
    assert lib.b2_voxel_downsample(h, pts.data_ptr(), 100, -0.5, None, None, out.data_ptr(), None, C.byref(n)) == INVALID
    assert "voxel" in _err(lib, h)
    assert lib.b2_estimate_normals(h, pts.data_ptr(), 100, -1.0, 20, out.data_ptr()) == INVALID
    assert lib.b2_estimate_normals(h, pts.data_ptr(), 100, 0.04, 0, out.data_ptr()) == INVALID
    assert lib.b2_transform(h, pts.data_ptr(), 100, None, out.data_ptr()) == INVALID
    assert lib.b2_set_icp_driver(h, 9) == INVALID
    assert lib.b2_ingest_pointcloud2(h, None, 1, 10, 12, 0, 4, 8, 0, 1, None, 0, out.data_ptr(), C.byref(n)) == INVALID
    assert lib.b2_slam_results(h, 5, None) == INVALID


def test_icp_parameter_validation(raw, b2):
    eng, lib, h = raw
    pts = eng.pack(np.random.default_rng(1).normal(size=(200, 3)).astype(np.float32))
    res = b2.IcpResult()
    for kw, msg in ((dict(max_corr=0.0), "max_correspondence_distance"), (dict(max_corr=-1.0), "max_correspondence_distance"),
                    (dict(max_iter=-3), "max_iteration"), (dict(mode=7), "mode")):
        p = b2.make_params(kw.get("max_corr", 0.05), kw.get("max_iter", 5))
        if "mode" in kw:
            p.mode = kw["mode"]
        rc = lib.b2_icp(h, pts.data_ptr(), 200, pts.data_ptr(), 200, None, None, C.byref(p), C.byref(res), None)
        assert rc == INVALID and msg in _err(lib, h), (kw, _err(lib, h))
    assert lib.b2_icp(h, pts.data_ptr(), 200, pts.data_ptr(), 200, None, None, None, C.byref(res), None) == INVALID
    p2l = b2.make_params(0.05, 5, mode=b2.P2L)  # point-to-plane without normals
    assert lib.b2_icp(h, pts.data_ptr(), 200, pts.data_ptr(), 200, None, None, C.byref(p2l), C.byref(res), None) == INVALID


def test_call_order_and_map_limits(raw, b2):
    eng, lib, h = raw
    pts = eng.pack(np.random.default_rng(2).normal(size=(300, 3)).astype(np.float32))
    res = b2.IcpResult()
    p = b2.make_params(0.05, 5)
    size = C.c_longlong(0)
    assert lib.b2_map_fuse(h, pts.data_ptr(), 300, None) == STATE           # no map yet
    assert lib.b2_map_size(h, C.byref(size)) == STATE
    assert lib.b2_slam_scan_dev(h, pts.data_ptr(), 300, 0.02, C.byref(p), C.byref(res)) == STATE
    assert "b2_map_init" in _err(lib, h)
    origin = (C.c_double * 3)(-5, -5, -5)
    assert lib.b2_map_init(h, 0.0, origin, 0.05, 1 << 16) == INVALID     # voxel must be > 0
    assert lib.b2_map_init(h, 0.05, origin, 0.05, 1 << 16) == 0
    wide = b2.make_params(0.5, 5)                                          # gate wider than the map's cell edge
    assert lib.b2_slam_scan_dev(h, pts.data_ptr(), 300, 0.02, C.byref(wide), C.byref(res)) == INVALID
    assert "exceeds" in _err(lib, h)
    p2l = b2.make_params(0.05, 5, mode=b2.P2L)                             # the map holds no normals
    assert lib.b2_slam_scan_dev(h, pts.data_ptr(), 300, 0.02, C.byref(p2l), C.byref(res)) == INVALID
    assert lib.b2_slam_scan_dev(h, pts.data_ptr(), 0, 0.02, C.byref(p), C.byref(res)) == INVALID  # empty scan
    assert lib.b2_slam_scan_dev(h, pts.data_ptr(), 300, 0.02, C.byref(p), C.byref(res)) == 0       # and a good one
    assert lib.b2_slam_results(h, 2, (b2.IcpResult * 2)()) == INVALID      # only one scan registered so far
    assert lib.b2_slam_results(h, 1, (b2.IcpResult * 1)()) == 0
    assert lib.b2_map_size(h, C.byref(size)) == 0 and size.value > 0


def test_scan_path_and_scan_downsample_arguments(raw, b2):
    """Round-2 entry points: b2_set_scan_path and b2_scan_downsample reject what they cannot serve; a scan with a
    non-finite coordinate has no voxel (Open3D's int32 voxel index is undefined for it): B2_ERR_KEY_RANGE."""
    eng, lib, h = raw
    assert lib.b2_set_scan_path(h, 7) == INVALID and "scan_path" in _err(lib, h)
    assert lib.b2_set_scan_path(h, 1) == 0 and lib.b2_set_scan_path(h, 0) == 0
    pts = eng.pack(np.random.default_rng(3).normal(size=(500, 3)).astype(np.float32))
    out = eng.empty_points(500)
    n = C.c_int(-1)
    assert lib.b2_scan_downsample(h, None, 500, 0.02, out.data_ptr(), C.byref(n)) == INVALID
    assert lib.b2_scan_downsample(h, pts.data_ptr(), 500, 0.0, out.data_ptr(), C.byref(n)) == INVALID
    assert lib.b2_scan_downsample(h, pts.data_ptr(), 500, 0.02, out.data_ptr(), None) == INVALID
    big = 300_000  # beyond the single-scan limit of the look-back scan (2048 blocks of 128 points)
    assert lib.b2_scan_downsample(h, pts.data_ptr(), big, 0.02, out.data_ptr(), C.byref(n)) == INVALID
    assert "b2_voxel_downsample" in _err(lib, h)
    assert lib.b2_scan_downsample(h, pts.data_ptr(), 0, 0.02, out.data_ptr(), C.byref(n)) == 0 and n.value == 0
    bad = np.random.default_rng(4).normal(size=(500, 3)).astype(np.float32)
    bad[123, 1] = np.nan
    assert lib.b2_scan_downsample(h, eng.pack(bad).data_ptr(), 500, 0.02, out.data_ptr(), C.byref(n)) == -4  # KEY_RANGE
    # the handle is usable again
    assert lib.b2_scan_downsample(h, pts.data_ptr(), 500, 0.02, out.data_ptr(), C.byref(n)) == 0 and 0 < n.value <= 500


def test_dense_driver_switch_and_correspondence_report_arguments(raw, b2):
    """b2_set_dense_driver / b2_map_icp_corr: unknown driver, NULL outputs, call before the map exists, and a slab map
    (voxel edge < gate), which has no per-query match report."""
    eng, lib, h = raw
    assert lib.b2_set_dense_driver(h, 5) == INVALID and "driver" in _err(lib, h)
    assert lib.b2_set_dense_driver(h, 1) == 0 and lib.b2_set_dense_driver(h, 0) == 0
    pts = eng.pack(np.random.default_rng(5).uniform(-1, 1, size=(400, 3)).astype(np.float32))
    prm, res = b2.make_params(0.05, 3), b2.IcpResult()
    d2 = eng.empty(400, torch.float64)
    match = eng.empty((400, 4), torch.float32)
    rc = lib.b2_map_icp_corr(h, pts.data_ptr(), 400, None, C.byref(prm), C.byref(res), d2.data_ptr(), match.data_ptr())
    assert rc == STATE  # no map yet
    eng.map_init(0.05, [-2.0, -2.0, -2.0], 0.05, 1 << 14)
    eng.map_fuse(pts)
    assert lib.b2_map_icp_corr(h, pts.data_ptr(), 400, None, C.byref(prm), C.byref(res), None, match.data_ptr()) == INVALID
    assert lib.b2_map_icp_corr(h, pts.data_ptr(), 0, None, C.byref(prm), C.byref(res), d2.data_ptr(), match.data_ptr()) == INVALID
    assert lib.b2_map_icp_corr(h, pts.data_ptr(), 400, None, C.byref(prm), C.byref(res), d2.data_ptr(), match.data_ptr()) == 0
    assert res.fitness > 0.9 and float(d2.max()) < 0.05 ** 2  # (nearly) every point is its own voxel's centroid
    eng.map_init(0.02, [-2.0, -2.0, -2.0], 0.05, 1 << 14)  # slab map
    eng.map_fuse(pts)
    rc = lib.b2_map_icp_corr(h, pts.data_ptr(), 400, None, C.byref(prm), C.byref(res), d2.data_ptr(), match.data_ptr())
    assert rc == INVALID and "brick-major" in _err(lib, h)
