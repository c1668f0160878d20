"""BASELINE.json-sized checks: 2 M-voxel resident map, 49k-pt scans.  The map itself is checked through
size-independent properties; the registration the bench times (dense-map ICP step kernel against the full
2 M-voxel table) is checked against the oracle run on the exported centroids."""
import numpy as np
import pytest
import torch

from conftest import assert_transform_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_map(engine):
    """~2 M voxels @ 5 cm: a 210 m corridor sampled densely (config 2's resident map)."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(7, length=210.0)
    engine.map_init(0.05, [-50.0, -50.0, -50.0], 0.05, 1 << 23)
    n_pts = 0
    for x0 in range(0, 210, 10):
        pts = synth.sample_scene_surface(scene, 0.02, (float(x0), float(x0 + 10)), seed=x0, device="cuda")
        p4 = torch.zeros((pts.shape[0], 4), dtype=torch.float32, device="cuda")
        p4[:, :3] = pts
        engine.map_fuse(p4)
        n_pts += pts.shape[0]
    engine.synchronize()
    return scene, n_pts


def test_full_size_map_occupancy_and_idempotence(engine, big_map):
    scene, n_pts = big_map
    n_vox = engine.map_size()
    assert 1_500_000 < n_vox < 3_000_000, n_vox
    cent, keys = engine.map_export()
    cnt = cent[:, 3].contiguous().view(torch.int32)
    assert int(cnt.sum().item()) == n_pts                      # every fused point is counted exactly once
    k = keys.cpu().numpy().astype(np.int64)
    packed = (k[:, 0] << 42) + (k[:, 1] << 21) + k[:, 2]
    assert (np.diff(packed) > 0).all()                         # canonical order, no duplicate voxel
    # every centroid lies inside its own voxel (convexity), so re-voxelising the centroids at the same
    # origin reproduces the key set (idempotence)
    c2, k2 = engine.voxel_downsample(cent.contiguous(), 0.05, origin=[-50.0, -50.0, -50.0])
    assert c2.shape[0] == cent.shape[0]
    assert torch.equal(k2, keys)


def test_full_size_scan_to_map_registration(engine, b2, big_map):
    """49k-pt scan against the 2 M-voxel map: the pose prior is perturbed and ICP must pull it back."""
    from lidar_slam_b200 import synth

    scene, _ = big_map
    pose = synth.look_at_pose([100.0, 1.0, 1.3], yaw=np.pi / 2 + 0.1, pitch=-0.1)
    scan = synth.render_scan(scene, pose, seed=9).numpy()
    sd, _ = engine.voxel_downsample(engine.pack(scan), 0.02)
    init = synth.perturb_pose(pose, 3, trans=0.02, rot_deg=0.3)
    res = engine.map_icp(sd, b2.make_params(0.05, 60), init=init)
    assert res.fitness > 0.9
    assert_transform_close(res.matrix(), pose, rad=4e-3, metres=1.5e-2)
    before = engine.map_size()
    engine.map_fuse(engine.pack(scan), res.matrix())
    after = engine.map_size()
    assert before <= after < before + 3000  # a registered scan lands almost entirely in existing voxels


def test_full_size_config2_registration_matches_the_oracle(engine, b2, big_map, dense_driver):
    """BASELINE config 2 at full size (icp_processor.py:103-113): three consecutive 49k-pt scans, each 2 cm
    down-sampled and registered against the 2 M-voxel map by the dense-map step kernel, versus
    ``registration_icp`` of the oracle on the map's exported centroids.  Iterations, correspondence count and
    fitness equal; transform within 1e-4 rad / 1e-4 m."""
    from lidar_slam_b200 import synth
    from oracle import cpu as orc

    scene, _ = big_map
    cent, _ = engine.map_export(want_keys=False)
    tgt = cent.cpu().numpy()[:, :3].astype(np.float64)
    assert len(tgt) > 1_500_000
    poses = synth.corridor_trajectory(scene, 3, start_x=60.0, step=0.02, seed=5)
    prm = b2.make_params(0.05, 30)
    for k, pose in enumerate(poses):
        scan = synth.render_scan(scene, pose, seed=700 + k).numpy()
        sd, _ = engine.voxel_downsample(engine.pack(scan), 0.02, want_keys=False)
        init = synth.perturb_pose(pose, 40 + k, trans=0.015, rot_deg=0.4)
        res = engine.map_icp(sd, prm, init=init)
        src = sd.cpu().numpy()[:, :3].astype(np.float64)
        ores = orc.registration_icp(src, tgt, 0.05, init=init, max_iteration=30)
        assert res.iterations == ores.iterations, k
        assert res.n_correspondences == int((ores.correspondence >= 0).sum()), k
        assert res.fitness == ores.fitness, k
        assert res.inlier_rmse == pytest.approx(ores.inlier_rmse, rel=1e-9, abs=1e-12)
        assert_transform_close(res.matrix(), ores.transformation)
        assert res.fitness > 0.8  # a real registration, not an empty overlap


def test_full_size_pair_through_the_sweep_driver_matches_the_oracle(engine, b2):
    """One full-size config-1 pair (49k-pt scans, 2 cm down-sampled) forced through the thread-per-query sweep
    kernel versus the oracle (round 1 compared the sweep driver with the group kernel only at this size)."""
    from lidar_slam_b200 import synth
    from oracle import cpu as orc

    src, tgt, _ = synth.make_pair(31)
    sd, _ = engine.voxel_downsample(engine.pack(src.numpy()), 0.02, want_keys=False)
    td, _ = engine.voxel_downsample(engine.pack(tgt.numpy()), 0.02, want_keys=False)
    engine.set_icp_driver("sweep")
    try:
        res, corr = engine.icp(sd, td, b2.make_params(0.05, 30), want_corr=True)
    finally:
        engine.set_icp_driver("auto")
    ores = orc.registration_icp(sd.cpu().numpy()[:, :3].astype(np.float64), td.cpu().numpy()[:, :3].astype(np.float64),
                                0.05, max_iteration=30)
    assert res.iterations == ores.iterations
    assert res.fitness == ores.fitness
    np.testing.assert_array_equal(corr.cpu().numpy(), ores.correspondence)
    assert_transform_close(res.matrix(), ores.transformation)


def test_large_batch_takes_the_sweep_driver_and_agrees_with_single_pair_runs(engine, b2):
    """16 full-size pairs (~470 k queries) make the automatic driver choice pick the thread-per-query sweep
    kernel with several CTAs per pair.  Size-independent properties: duplicated pairs give bit-identical
    records wherever they sit in the batch, and every record equals the single-pair registration of the same
    clouds (which runs the 4-lane-group kernel): same iterations, same fitness, transform within tolerance."""
    import torch
    from conftest import assert_transform_close
    from lidar_slam_b200 import synth

    distinct = [synth.make_pair(s) for s in (20, 21, 22, 23)]
    order = [0, 1, 2, 3, 3, 2, 1, 0, 0, 0, 1, 1, 2, 2, 3, 3]
    srcs = [distinct[i][0].numpy() for i in order]
    tgts = [distinct[i][1].numpy() for i in order]
    n = len(srcs[0])
    offs = torch.arange(0, len(order) + 1, dtype=torch.int32, device="cuda") * n
    prm = b2.make_params(0.05, 30)
    res = engine.icp_batch(engine.pack(np.concatenate(srcs)), offs, engine.pack(np.concatenate(tgts)), offs, prm,
                           ds_voxel=0.02)
    engine.synchronize()
    rec = b2.results_to_numpy(res)
    first = {}
    for pos, i in enumerate(order):
        if i in first:
            assert rec[pos].tobytes() == rec[first[i]].tobytes(), (pos, i)
        else:
            first[i] = pos
    for i, pos in first.items():
        sd, _ = engine.voxel_downsample(engine.pack(srcs[pos]), 0.02, want_keys=False)
        td, _ = engine.voxel_downsample(engine.pack(tgts[pos]), 0.02, want_keys=False)
        single, _ = engine.icp(sd, td, prm)
        assert rec["iterations"][pos] == single.iterations
        assert rec["fitness"][pos] == single.fitness
        assert rec["n_correspondences"][pos] == single.n_correspondences
        assert_transform_close(rec["transformation"][pos], single.matrix())


def test_full_size_streaming_steps_match_the_resident_cpu_map(b2):
    """The path bench.py times, end to end at full size: the ~2.1 M-voxel bench map, then consecutive 49k-pt
    bench scans through b2_slam_scan (eager, captured and graph-replayed steps: down-sample -> dense-map ICP ->
    fuse), against the resident CPU map of the oracle (oracle.cpp ResidentMap, the twin of
    pipeline.ResidentSequence) fed the same samples and scans (icp_processor.py:43-119).  Iterations and
    fitness equal, poses within 1e-4 rad / 1e-4 m, and after the sequence the two maps hold the same voxels
    with the same point counts and centroids within 1e-5."""
    import bench
    from oracle import cpu as orc

    eng = b2.Engine(0)
    scene = bench.build_scene(0)
    n_vox = bench.preload_map(eng, scene, "cuda")
    cpu_map = orc.ResidentMapCPU(bench.MAP_VOXEL, bench.MAP_ORIGIN, ds_voxel=bench.DS_VOXEL, max_corr=bench.MAX_CORR,
                                 max_iteration=bench.MAX_ITER)
    for pts in bench.map_chunks(scene, "cpu"):
        cpu_map.fuse(pts.numpy())
    assert cpu_map.size() == n_vox
    scans, poses = bench.make_scans(scene, 6, "cuda", seed=1)
    prm = b2.make_params(bench.MAX_CORR, bench.MAX_ITER)
    eng.set_pose(poses[0])
    cpu_map.pose = poses[0]
    for k, s in enumerate(scans):
        res = eng.slam_scan(s, bench.DS_VOXEL, prm)
        ores = cpu_map.process(s)
        assert res.iterations == ores.iterations, k
        assert res.fitness == ores.fitness, k
        assert res.inlier_rmse == pytest.approx(ores.inlier_rmse, rel=1e-9, abs=1e-12)
        assert_transform_close(res.matrix(), ores.transformation)
        assert res.fitness > 0.9
    cent, keys = eng.map_export()
    oc, ok, on = cpu_map.map_points()
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    c = cent.cpu().numpy()
    np.testing.assert_array_equal(c[:, 3].copy().view(np.int32), on)
    np.testing.assert_allclose(c[:, :3], oc, rtol=1e-5, atol=1e-5)
    cpu_map.close()
    eng.close()
