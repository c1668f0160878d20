"""Round-2 regression tests through the C ABI: the deferred-scan count ring, the widened sort budget of the
scan step, asynchronous batch status, and the count-weighted partial-map merge."""
import numpy as np
import pytest
import torch

from oracle import cpu as orc

pytestmark = pytest.mark.gpu


def f64(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def test_deferred_scans_of_distinct_sizes_equal_synchronous_results(b2):
    """48 scans with a Bernoulli pixel drop (every scan a different point count, as the phone sends them:
    ARDepthViewController.swift:112) submitted with out == NULL — the host runs far ahead of the device —
    must give the records of the scan-by-scan synchronous run bit for bit.  Round 1 staged the live count in
    an 8-slot pinned ring that a run-ahead of more than 8 scans overwrote."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(3, length=14.0)
    poses = synth.corridor_trajectory(scene, 48, seed=3, step=0.02)
    scans = [synth.project_records(synth.scan_records(synth.render_depth(scene, p, seed=300 + k), keep_prob=0.5,
                                                      seed=k)).numpy().astype(np.float32) for k, p in enumerate(poses)]
    assert len({len(s) for s in scans}) > 40  # sizes really differ
    prm = b2.make_params(0.05, 30)
    out = []
    for deferred in (False, True):
        eng = b2.Engine(0)
        eng.map_init(0.05, [-20.0, -20.0, -20.0], 0.05, 1 << 19)
        pinned = [torch.from_numpy(s).pin_memory() for s in scans]
        if deferred:
            for p in pinned:
                eng.slam_scan(p.numpy(), 0.02, prm, want_result=False)
            res = eng.slam_results(len(scans))
        else:
            res = [eng.slam_scan(p.numpy(), 0.02, prm) for p in pinned]
        cent, keys = eng.map_export()
        out.append((res, cent.cpu().numpy().copy(), keys.cpu().numpy().copy()))
        eng.close()
    (ra, ca, ka), (rb, cb, kb) = out
    for k in range(1, len(scans)):
        assert bytes(ra[k]) == bytes(rb[k]), k
    np.testing.assert_array_equal(ka, kb)
    np.testing.assert_array_equal(ca, cb)


def _wide_scans(seed, n=3):
    """Scans whose voxel extent needs more than the 4 radix passes (32 key bits) the scan step launches by
    default: 60 m across at 2 cm voxels = 12 bits per axis."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(-30.0, 30.0, size=(30000, 3)).astype(np.float32)
    return [base + np.float32(0.001 * k) for k in range(n)]


def test_scan_step_widens_its_sort_budget_instead_of_failing(b2):
    scans = _wide_scans(0)
    eng = b2.Engine(0)
    origin = [-100.0, -100.0, -100.0]
    eng.map_init(0.05, origin, 0.05, 1 << 23)  # (random points: one brick of the pool per point)
    eng.set_scan_path("sorted")          # (the default scratch-hash passes have no sort budget to exceed)
    prm = b2.make_params(0.05, 5)
    poses = []
    for s in scans + scans:              # first scan: fuse only; then eager, captured and replayed steps
        res = eng.slam_scan(s, 0.02, prm)
        poses.append(res.matrix())
    assert res.fitness > 0.9             # the last scans register against their own earlier copies
    _, keys = eng.map_export()
    want = set()
    for s, T in zip(scans + scans, poses):
        want |= {tuple(k) for k in orc.voxel_keys(orc.transform(f64(s), T), 0.05, origin=np.array(origin)).tolist()}
    got = {tuple(k) for k in keys.cpu().numpy().tolist()}
    assert len(got ^ want) <= 3          # (a point within rounding of a voxel face may flip)
    eng.close()


def test_deferred_scan_failure_leaves_pose_and_map_untouched_and_can_be_resubmitted(b2):
    scans = _wide_scans(1)
    eng = b2.Engine(0)
    eng.map_init(0.05, [-100.0, -100.0, -100.0], 0.05, 1 << 23)
    eng.set_scan_path("sorted")
    prm = b2.make_params(0.05, 5)
    pinned = [torch.from_numpy(s).pin_memory().numpy() for s in scans]
    for p in pinned:
        eng.slam_scan(p, 0.02, prm, want_result=False)
    with pytest.raises(b2.B2Error, match="resubmit") as ei:
        eng.slam_results(len(scans))
    assert ei.value.code == -4           # B2_ERR_KEY_RANGE
    assert eng.map_size() == 0           # nothing was committed
    np.testing.assert_array_equal(eng.get_pose(), np.eye(4))
    for p in pinned:                     # the budget has been widened: the same scans now go through
        eng.slam_scan(p, 0.02, prm, want_result=False)
    res = eng.slam_results(len(scans))
    assert res[-1].fitness > 0.9 and eng.map_size() > 20000
    eng.close()


def test_batch_reports_device_failures_in_the_call_that_caused_them(engine, b2):
    """A batch whose targets span more than the 16-bit cell range of a pair grid fails on the device; with the
    default sync=True the failure is raised by icp_batch itself, not by the next unrelated call."""
    rng = np.random.default_rng(5)
    far = np.concatenate([rng.uniform(0, 1, (500, 3)), rng.uniform(0, 1, (500, 3)) + 5000.0]).astype(np.float32)
    near = rng.uniform(0, 1, (1000, 3)).astype(np.float32)
    offs = torch.tensor([0, 1000], dtype=torch.int32, device="cuda")
    with pytest.raises(b2.B2Error):
        engine.icp_batch(engine.pack(near), offs, engine.pack(far), offs, b2.make_params(0.05, 5))
    # and the handle is usable again right away
    res = engine.icp_batch(engine.pack(near), offs, engine.pack(near), offs, b2.make_params(0.05, 5))
    rec = b2.results_to_numpy(res)
    assert rec["fitness"][0] == 1.0


def test_partial_maps_merge_to_the_voxel_downsample_of_the_concatenation(b2):
    """SURVEY.md 8(e) / icp_processor.py:74,132-133: two ranks each fuse half of a trajectory; the gathered
    (key, sum xyz, count) records merged with the count-weighted centroid merge must equal voxel_down_sample
    of ALL points: keys and counts bit-exact, centroids within 1e-5."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(4, length=12.0)
    poses = synth.corridor_trajectory(scene, 6, seed=4, step=0.3)
    scans = [synth.render_scan(scene, p, seed=400 + k).numpy() for k, p in enumerate(poses)]
    origin = np.array([-20.0, -20.0, -20.0])
    halves = []
    for part in (range(0, 3), range(3, 6)):
        eng = b2.Engine(0)
        eng.map_init(0.05, origin, 0.05, 1 << 18)
        for k in part:
            eng.map_fuse(eng.pack(scans[k]), poses[k])
        keys, sums = eng.map_export_records()
        halves.append((keys.clone(), sums.clone()))
        eng.close()
    merged = b2.Engine(0)
    merged.map_init(0.05, origin, 0.05, 1 << 18)
    for keys, sums in halves:
        merged.map_merge_records(keys, sums)
    cent, keys = merged.map_export()
    world = np.concatenate([orc.transform(f64(s), p) for s, p in zip(scans, poses)])
    # the GPU stores the transformed points' sums in fp64 without an intermediate float32 rounding, like this
    oc, ok, on = orc.voxel_downsample(world, 0.05, origin=origin)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    np.testing.assert_array_equal(cent.cpu().numpy()[:, 3].copy().view(np.int32), on)
    np.testing.assert_allclose(cent.cpu().numpy()[:, :3], oc, rtol=1e-5, atol=1e-5)
    # merging the same records into a map that already holds the first half adds up sums and counts
    k0, s0 = halves[0]
    merged.map_merge_records(k0, s0)
    _, s2 = merged.map_export_records()
    assert float(s2[:, 3].sum().item()) == float(on.sum()) + float(s0[:, 3].sum().item())
    merged.close()


def _rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_scan_downsample_without_a_sort_is_bit_identical_to_the_oracle(engine):
    """b2_scan_downsample (scratch hash + fp64 atomics, what b2_slam_scan runs in front of ICP,
    icp_processor.py:90): the sum of the float32 coordinates of one voxel is exact in fp64, so the centroids
    and counts equal the oracle's in-order sums bit for bit; rows come in order of first appearance."""
    from lidar_slam_b200 import synth

    for seed, keep in ((0, 1.0), (1, 0.6)):
        scene = synth.make_corridor_scene(seed, length=14.0)
        pose = synth.corridor_trajectory(scene, 1, seed=seed)[0]
        scan = synth.project_records(synth.scan_records(synth.render_depth(scene, pose, seed=seed), keep_prob=keep,
                                                        seed=seed)).numpy()
        got = engine.scan_downsample(engine.pack(scan), 0.02).cpu().numpy()
        oc, ok, on = orc.voxel_downsample(f64(scan), 0.02)
        assert len(got) == len(oc)
        want = np.concatenate([oc.astype(np.float32), on.view(np.float32)[:, None]], axis=1)
        np.testing.assert_array_equal(_rows_sorted(got).view(np.int32), _rows_sorted(want).view(np.int32))
        # order of first appearance: the voxel of input point 0 comes first, and the row order follows the
        # first input index of every voxel
        keys = orc.voxel_keys(f64(scan), 0.02)
        first = {}
        for i, k in enumerate(map(tuple, keys.tolist())):
            first.setdefault(k, i)
        order = sorted(first, key=first.get)
        pos = {tuple(k): j for j, k in enumerate(ok.tolist())}
        want_rows = want[[pos[k] for k in order]]
        np.testing.assert_array_equal(got.view(np.int32), want_rows.view(np.int32))


def test_hash_and_sorted_scan_paths_register_and_fuse_alike(b2):
    """The sort-free scan step against the radix-sort scan step on the same sequence (eager, captured and
    replayed steps): same iterations and fitness, poses within 1e-7 (the source order differs, hence the summation order), the same voxels with the same point counts
    in the map, centroids within 1e-6 relative (per-scan sums: fixed point vs in-order fp64)."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(11, length=14.0)
    poses = synth.corridor_trajectory(scene, 8, seed=11, step=0.02)
    scans = [synth.render_scan(scene, p, seed=900 + k).numpy() for k, p in enumerate(poses)]
    out = {}
    for path in ("hash", "sorted"):
        eng = b2.Engine(0)
        eng.map_init(0.05, [-20.0, -20.0, -20.0], 0.05, 1 << 19)
        eng.set_scan_path(path)
        res = [eng.slam_scan(s, 0.02, b2.make_params(0.05, 30)) for s in scans]
        cent, keys = eng.map_export()
        out[path] = (res, cent.cpu().numpy().copy(), keys.cpu().numpy().copy(), eng.launch_count())
        eng.close()
    (ra, ca, ka, la), (rb, cb, kb, lb) = out["hash"], out["sorted"]
    assert la < lb - 100                 # really a different, shorter pipeline
    for k in range(1, len(scans)):
        assert ra[k].iterations == rb[k].iterations > 0, k
        assert ra[k].fitness == rb[k].fitness, k
        np.testing.assert_allclose(ra[k].matrix(), rb[k].matrix(), rtol=0, atol=1e-7)
    np.testing.assert_array_equal(ka, kb)
    np.testing.assert_array_equal(ca[:, 3].copy().view(np.int32), cb[:, 3].copy().view(np.int32))
    np.testing.assert_allclose(ca[:, :3], cb[:, :3], rtol=1e-6, atol=1e-6)


def test_hash_scan_path_on_a_slab_map_matches_the_resident_cpu_map(b2):
    """The reference's default geometry (2 cm voxels, 5 cm gate: a map of 3^3-voxel cells, the slab layout and the
    group ICP kernel) through the sort-free scan step, against the resident CPU map of the oracle fed the same scans
    (icp_processor.py:43-119): iterations and fitness equal, poses within 1e-4, identical voxel keys and counts."""
    from conftest import assert_transform_close
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(12, length=12.0)
    poses = synth.corridor_trajectory(scene, 5, seed=12, step=0.02)
    scans = [synth.render_scan(scene, p, seed=40 + k).numpy()[::2].copy() for k, p in enumerate(poses)]
    origin = [-20.0, -20.0, -20.0]
    eng = b2.Engine(0)
    eng.map_init(0.02, origin, 0.05, 1 << 19)
    cpu_map = orc.ResidentMapCPU(0.02, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30)
    prm = b2.make_params(0.05, 30)
    for k, s in enumerate(scans):
        res = eng.slam_scan(s, 0.02, prm)
        ores = cpu_map.process(s)
        if k == 0:
            assert ores is None
            continue
        assert res.iterations == ores.iterations and res.fitness == ores.fitness, k
        assert_transform_close(res.matrix(), ores.transformation)
    cent, keys = eng.map_export()
    oc, ok, on = cpu_map.map_points()
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    c = cent.cpu().numpy()
    np.testing.assert_array_equal(c[:, 3].copy().view(np.int32), on)
    np.testing.assert_allclose(c[:, :3], oc, rtol=1e-5, atol=1e-5)
    cpu_map.close()
    eng.close()


@pytest.mark.parametrize("path", ["hash", "sorted"])
def test_scan_step_edge_sizes_and_no_downsample(b2, path):
    """Tiny and ragged scans through the scan step (1, 2, 31, 33, 1025 points; ds_voxel = 0: the scan is the ICP
    source as it is), eager / captured / replayed: both scan paths keep the map equal to the oracle's voxel
    down-sample of everything fused (keys and counts exact), and a one-point scan is a legal scan."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(21, length=8.0)
    pose = synth.corridor_trajectory(scene, 1, seed=21)[0]
    full = synth.render_scan(scene, pose, seed=5).numpy()
    origin = np.array([-20.0, -20.0, -20.0])
    eng = b2.Engine(0)
    eng.map_init(0.05, origin, 0.05, 1 << 18)
    eng.set_scan_path(path)
    prm = b2.make_params(0.05, 10)
    fused = []
    # the first scan is the map; the others (subsets of it) register against its voxel centroids
    for n, ds in ((len(full), 0.02), (1, 0.02), (2, 0.0), (31, 0.02), (33, 0.0), (1025, 0.02), (1025, 0.02), (1025, 0.02)):
        s = full[:: max(1, len(full) // n)][:n].copy()
        eng.set_pose(np.eye(4))              # every scan starts from the identity prior
        res = eng.slam_scan(s, ds, prm)
        fused.append((s, res.matrix()))
        if len(fused) > 1:
            assert np.isfinite(res.matrix()).all() and 0.0 <= res.fitness <= 1.0 and res.iterations <= 10
            if n > 2:    # (one or two correspondences leave the rotation undetermined: the SVD fall-back's choice)
                assert np.abs(res.matrix() - np.eye(4)).max() < 0.05   # a subset of the map: the pose stays put
    _, keys = eng.map_export()
    cent, _ = eng.map_export(want_keys=False)
    world = np.concatenate([orc.transform(f64(s), T) for s, T in fused])
    _, ok, on = orc.voxel_downsample(world, 0.05, origin=origin)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    np.testing.assert_array_equal(cent.cpu().numpy()[:, 3].copy().view(np.int32), on)
    eng.close()
