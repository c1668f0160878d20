"""The B200 processor behind the reference's plugin interface, against the reference sequence
restated on the CPU oracle (oracle/pipeline.py)."""
import multiprocessing
import os

import numpy as np
import pytest

from conftest import assert_transform_close
from oracle import pipeline

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, "3d-lidar-slam_b200", "config", "lidar_config.yaml")


def _scans(n, seed=2):
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(seed, length=10.0)
    poses = synth.corridor_trajectory(scene, n, seed=seed, step=0.02)
    return [synth.render_scan(scene, p, seed=50 + k).numpy() for k, p in enumerate(poses)]


def _processor(**kw):
    from lidar_slam_b200.plugin import DataTransfer, set_algorithm

    dt = DataTransfer()
    return set_algorithm("b200_icp", CONFIG, dt, multiprocessing.Event(), **kw), dt


def test_reference_mode_follows_the_reference_sequence():
    """icp_processor.py:43-119 step by step: per-cloud voxel origin, raw map, scan rows first."""
    scans = _scans(4)
    proc, _ = _processor(mode="reference", max_iteration=60)
    # teacher-forced twin: the oracle rounds every stored cloud to float32 like the GPU's HBM layout
    ref = pipeline.ReferenceSequence(max_iteration=60, maintenance="off", f32_storage=True)
    # pure float64 sequence (what Open3D holds in memory): same poses within the stated tolerance;
    # the 1e-6 convergence test may fire one iteration apart
    ref64 = pipeline.ReferenceSequence(max_iteration=60, maintenance="off")
    for k, s in enumerate(scans):
        proc.construct_global_map(s)
        proc.point_clouds_in_map += 1  # what ProcessPointClouds.process does after the call
        ref.process(s)
        ref64.process(s)
        assert proc.point_clouds_in_map == ref.point_clouds_in_map
        if k == 0:
            continue
        assert proc.last_result.iterations == ref.last_result.iterations
        assert_transform_close(proc.previous_transformation[-1], ref.previous_transformation[-1])
        assert abs(proc.last_result.iterations - ref64.last_result.iterations) <= 2
        assert_transform_close(proc.previous_transformation[-1], ref64.previous_transformation[-1],
                               rad=1e-4, metres=2e-4)
    gm = proc.global_map
    assert gm.shape == ref.global_map.shape == (4 * 49152, 3)
    np.testing.assert_allclose(gm, ref.global_map, rtol=1e-5, atol=1e-4)
    np.testing.assert_array_equal(gm[-49152:], scans[0])  # first scan is the map frame, kept verbatim at the end
    assert len(proc.previous_transformation) == 4


def test_resident_mode_through_the_queue():
    """process(): queue -> projection -> construct_global_map, first-scan double increment
    (icp_processor.py:54 + generic_point_cloud_processor.py:126), lazy global_map, reset()."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(4, length=10.0)
    poses = synth.corridor_trajectory(scene, 3, seed=4, step=0.02)
    proc, dt = _processor(mode="resident", max_iteration=30, map_voxel=0.05)
    for k, p in enumerate(poses):
        rec = synth.scan_records(synth.render_depth(scene, p, seed=70 + k)).numpy()
        dt.pixel_depth_map_queue.put(rec)
        proc.process()
    assert proc.point_clouds_in_map == 4  # 3 scans + the reference's extra increment on the first
    assert len(proc.previous_transformation) == 3
    assert proc.last_result.fitness > 0.5
    gm = proc.global_map
    assert gm.ndim == 2 and gm.shape[1] == 3 and len(gm) > 3000 and gm.dtype == np.float32
    assert proc.downsample_global_map() is gm
    # the pose chain moves along the trajectory (constant-pose prior, icp_processor.py:106)
    step = np.linalg.norm(proc.previous_transformation[-1][:3, 3] - proc.previous_transformation[-2][:3, 3])
    assert 0.0 < step < 0.1
    proc.reset_event.set()
    proc.reset()
    assert proc.global_map is None and proc.point_clouds_in_map == 0 and not proc.reset_event.is_set()
    np.testing.assert_array_equal(proc.previous_transformation[-1], np.eye(4))


def test_reference_mode_periodic_revoxelisation():
    scans = _scans(2)
    proc, _ = _processor(mode="reference", max_iteration=20, maintenance="voxel")
    proc.construct_global_map(scans[0])
    proc.point_clouds_in_map = 10  # next scan hits the count % 10 == 0 cadence (icp_processor.py:127-133)
    proc.construct_global_map(scans[1])
    n = len(proc.global_map)
    assert 0 < n <= 2 * 49152
    from oracle import cpu as orc

    ref = pipeline.ReferenceSequence(max_iteration=20, maintenance="voxel")
    ref.process(scans[0])
    ref.point_clouds_in_map = 10
    ref.construct_global_map(scans[1])
    assert abs(n - len(ref.global_map)) <= 2  # 1 mm voxels: float32 vs float64 storage can split a face case
    del orc


def test_streaming_replay_tracks_and_grows_the_map():
    """BASELINE config 4 in miniature: 30 consecutive frames through the plugin (resident mode, reference
    defaults incl. the 2 cm map voxel -> cells of 3x3x3 voxels); the pose chain must follow the ground truth
    and the map must keep growing."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(11, length=14.0)
    poses = synth.corridor_trajectory(scene, 30, start_x=3.0, step=0.025, seed=3)
    proc, _ = _processor(mode="resident", max_iteration=30)
    inv0 = np.linalg.inv(poses[0])
    sizes = []
    for k, p in enumerate(poses):
        proc.construct_global_map(synth.render_scan(scene, p, seed=200 + k).numpy())
        proc.point_clouds_in_map += 1
        sizes.append(proc._engine.map_size())
        if k:
            T_true = inv0 @ poses[k]
            err = np.linalg.norm(proc.previous_transformation[-1][:3, 3] - T_true[:3, 3])
            # point-to-point ICP capped at 30 iterations lags a little behind a moving sensor and the map is
            # built from its own estimates, so the drift bound grows with the frame index
            assert err < 0.03 + 0.01 * k, (k, err)
            assert proc.last_result.fitness > 0.9
    assert all(b >= a for a, b in zip(sizes, sizes[1:])) and sizes[-1] > 1.3 * sizes[0]
    assert len(proc.global_map) == sizes[-1]


@pytest.mark.parametrize("count", [10, 40])
def test_reference_mode_full_maintenance(count):
    """do_stuff with the outlier filters (icp_processor.py:135-148) against the oracle sequence."""
    scans = [s[::3].copy() for s in _scans(2)]  # thinned: the CPU oracle's 80-NN pass stays in seconds
    proc, _ = _processor(mode="reference", max_iteration=20, maintenance="full")
    ref = pipeline.ReferenceSequence(max_iteration=20, maintenance="full", f32_storage=True)
    proc.construct_global_map(scans[0])
    ref.process(scans[0])
    proc.point_clouds_in_map = ref.point_clouds_in_map = count
    proc.construct_global_map(scans[1])
    ref.construct_global_map(scans[1])
    n, n_ref = len(proc.global_map), len(ref.global_map)
    assert 0 < n < 2 * len(scans[0])
    assert abs(n - n_ref) <= max(3, n_ref // 2000), (n, n_ref)  # threshold-straddling points only


def test_process_with_fused_projection_equals_host_projection():
    """process() (records -> device projection -> registration) == construct_global_map(project_pixel_to_3d(records))."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(6, length=10.0)
    poses = synth.corridor_trajectory(scene, 3, seed=6, step=0.02)
    recs = [synth.scan_records(synth.render_depth(scene, p, seed=90 + k), keep_prob=0.6, seed=k).numpy()
            for k, p in enumerate(poses)]
    a, dta = _processor(mode="resident", max_iteration=30, map_voxel=0.05)
    b, _ = _processor(mode="resident", max_iteration=30, map_voxel=0.05)
    for r in recs:
        dta.pixel_depth_map_queue.put(r)
        a.process()
        b.construct_global_map(b.project_pixel_to_3d(r))
        b.point_clouds_in_map += 1
    assert a.point_clouds_in_map == b.point_clouds_in_map == 4
    np.testing.assert_array_equal(a.previous_transformation[-1], b.previous_transformation[-1])
    np.testing.assert_array_equal(a.global_map, b.global_map)


def test_process_loop_batches_scans_through_the_real_processor():
    """ProcessLoop (row f4) drives the B200 processor: same poses and map as scan-by-scan process()."""
    import threading
    import time

    from lidar_slam_b200 import synth
    from lidar_slam_b200.plugin.process_pc_manager import ProcessLoop

    scene = synth.make_corridor_scene(6, length=10.0)
    poses = synth.corridor_trajectory(scene, 5, seed=6, step=0.02)
    recs = [synth.scan_records(synth.render_depth(scene, p, seed=90 + k)).numpy() for k, p in enumerate(poses)]
    a, dta = _processor(mode="resident", max_iteration=30, map_voxel=0.05)
    b, dtb = _processor(mode="resident", max_iteration=30, map_voxel=0.05)
    for r in recs:
        dta.pixel_depth_map_queue.put(r)
        dtb.pixel_depth_map_queue.put(r)
    stop, reset = multiprocessing.Event(), multiprocessing.Event()
    # 590 KB records cross the queue's 64 KB pipe through a feeder thread: the batch forms because the
    # dequeue lingers for the next record (a generous linger keeps the assertion below load-independent)
    loop = ProcessLoop(a, dta, stop, reset, max_batch=4, idle_timeout=0.05, linger=0.5)
    th = threading.Thread(target=loop.run, daemon=True)
    th.start()
    for _ in recs:
        b.process()
    t0 = time.time()
    while loop.scans_processed < len(recs) and time.time() - t0 < 60:
        time.sleep(0.01)
    stop.set()
    th.join(timeout=10)
    assert loop.scans_processed == len(recs)
    assert loop.batches == 2, loop.batches  # 5 queued scans, max_batch 4 -> batches of 4 + 1
    np.testing.assert_array_equal(a.previous_transformation[-1], b.previous_transformation[-1])
    published = dta.global_map_queue.get(timeout=5)
    assert published.shape[1] == 3 and len(published) > 1000
    np.testing.assert_array_equal(a.global_map, b.global_map)


def test_resident_mode_is_asynchronous_and_synchronises_on_observation():
    """Resident mode only enqueues a scan: the poses arrive when previous_transformation /
    last_result / global_map are read.  Whatever the read pattern, the results are those of the scan-by-scan
    synchronous C-ABI run, bit for bit (icp_processor.py:106,118: the reference reads the list only inside ICP)."""
    import lidar_slam_b200 as b2

    scans = _scans(12, seed=6)
    eng = b2.Engine(0)
    eng.map_init(0.02, [0.0, 0.0, 0.0], 0.05, 1 << 20)
    prm = b2.make_params(0.05, 30)
    want = [eng.slam_scan(s, 0.02, prm).matrix() for s in scans]
    eng.close()
    for read_every in (1, 5, 100):
        proc, _ = _processor(mode="resident", max_iteration=30, map_capacity_cells=1 << 20)
        for k, s in enumerate(scans):
            proc.construct_global_map(s)
            proc.point_clouds_in_map += 1
            if (k + 1) % read_every == 0:
                assert len(proc.previous_transformation) == k + 1      # identity + k registrations (first scan: none)
        assert len(proc._pending_first) == (0 if read_every == 1 else len(scans) % read_every)
        assert proc.last_result is not None and proc.last_result.fitness > 0.9   # (reading flushes)
        assert not proc._pending_first
        got = list(proc.previous_transformation)
        assert len(got) == len(scans)
        np.testing.assert_array_equal(got[0], np.eye(4))
        for k in range(1, len(scans)):
            np.testing.assert_array_equal(got[k], want[k])
        assert proc.global_map.shape[1] == 3 and len(proc.global_map) > 1000
        proc.reset()
        assert len(proc.previous_transformation) == 1 and proc.last_result is None and proc.point_clouds_in_map == 0
        proc.engine.close()


def test_pageable_scans_are_copied_before_the_call_returns():
    """b2_slam_scan stages a pageable array through the library's pinned ring: the caller may overwrite or free
    it as soon as the call returns, even though the scan has only been enqueued.  Same results as pinned input."""
    import torch

    scans = _scans(10, seed=7)
    runs = []
    for kind in ("pinned", "pageable-clobbered"):
        proc, _ = _processor(mode="resident", max_iteration=30, map_capacity_cells=1 << 20)
        keep = []
        for s in scans:
            if kind == "pinned":
                a = torch.from_numpy(s.copy()).pin_memory().numpy()
                keep.append(a)                       # a pinned array is read in place: keep it alive and unchanged
                proc.construct_global_map(a)
            else:
                a = s.copy()
                proc.construct_global_map(a)
                a[:] = np.nan                        # clobber right away
            proc.point_clouds_in_map += 1
        runs.append([T.copy() for T in proc.previous_transformation])
        proc.engine.close()
    assert len(runs[0]) == len(runs[1]) == len(scans)
    for a, b in zip(*runs):
        np.testing.assert_array_equal(a, b)


def test_deferred_errors_surface_at_the_read_and_leave_the_processor_usable():
    """A scan the device cannot voxelise (a NaN coordinate) fails asynchronously: construct_global_map returns,
    the B2Error is raised where the poses are read, nothing of that scan is committed, and later scans work."""
    import lidar_slam_b200 as b2

    scans = _scans(4, seed=8)
    proc, _ = _processor(mode="resident", max_iteration=30, map_capacity_cells=1 << 20)
    proc.construct_global_map(scans[0]); proc.point_clouds_in_map += 1
    proc.construct_global_map(scans[1]); proc.point_clouds_in_map += 1
    assert len(proc.previous_transformation) == 2
    size_before = proc.engine.map_size()
    bad = scans[2].copy()
    bad[17, 2] = np.nan
    proc.construct_global_map(bad)                   # enqueued; no error yet
    with pytest.raises(b2.B2Error):
        len(proc.previous_transformation)
    assert not proc._pending_first                   # not retried against later scans' records
    assert proc.engine.map_size() == size_before     # the failed scan touched neither map ...
    np.testing.assert_array_equal(proc.engine.get_pose(), proc.previous_transformation[-1])  # ... nor pose
    proc.construct_global_map(scans[2]); proc.point_clouds_in_map += 1
    assert len(proc.previous_transformation) == 3 and proc.last_result.fitness > 0.9
    proc.engine.close()


def test_process_records_fuses_projection_into_the_scan_step():
    """process_records (the body of ProcessPointClouds.process) feeds the wire-format records straight to
    b2_slam_scan_records; the result must equal projecting on the host (generic_point_cloud_processor.py:93-108,
    the NumPy regime of this interpreter) and calling construct_global_map, bit for bit."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(9, length=10.0)
    poses = synth.corridor_trajectory(scene, 6, seed=9, step=0.02)
    recs = [synth.scan_records(synth.render_depth(scene, p, seed=80 + k), keep_prob=0.9, seed=k).numpy()
            for k, p in enumerate(poses)]
    a, _ = _processor(mode="resident", max_iteration=30, map_capacity_cells=1 << 20)
    b, _ = _processor(mode="resident", max_iteration=30, map_capacity_cells=1 << 20)
    for r in recs:
        a.process_records(r)
        b.construct_global_map(b.project_pixel_to_3d(r))
        b.point_clouds_in_map += 1
    assert a.point_clouds_in_map == b.point_clouds_in_map
    ta, tb = list(a.previous_transformation), list(b.previous_transformation)
    assert len(ta) == len(tb) == len(recs)
    for x, y in zip(ta, tb):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(a.global_map, b.global_map)
    a.engine.close()
    b.engine.close()
