"""GPU-vs-oracle parity through the C ABI (libb2slam.so via ctypes).

Bar (BASELINE.json north_star): voxel keys, voxel occupancy and correspondence indices
bit-exact; final 4x4 transforms within 1e-4 rad / 1e-4 m; voxel centroids within
1e-5 relative / 1e-4 m.  Stage tests are teacher-forced: the oracle is fed the very
same float32 coordinates the kernel reads (widened to float64, which is exact), so
integer and index outputs must match bit for bit.
"""
import numpy as np
import pytest
import torch

from conftest import assert_transform_close
from oracle import cpu as orc
from oracle import pipeline

pytestmark = pytest.mark.gpu

CENT_ATOL = 1e-4  # metres
CENT_RTOL = 1e-5


def f64(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def xyz(t):
    return t.cpu().numpy()[:, :3]


def counts(t):
    return t.cpu().numpy()[:, 3].copy().view(np.int32)


# ------------------------------------------------------------------ voxel down-sample (A.1)
@pytest.mark.parametrize("voxel", [0.02, 0.05, 0.3])
def test_voxel_downsample_bit_exact(engine, pair0, voxel):
    src, _, _ = pair0
    cent, keys = engine.voxel_downsample(engine.pack(src), voxel)
    oc, ok, on = orc.voxel_downsample(f64(src), voxel)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    np.testing.assert_array_equal(counts(cent), on)
    np.testing.assert_array_equal(xyz(cent), oc.astype(np.float32))  # fp64 sums in input order, rounded once


def test_voxel_downsample_explicit_origin_and_transform(engine, pair0):
    src, _, T = pair0
    origin = np.array([-7.013, -7.5, -6.25])
    cent, keys = engine.voxel_downsample(engine.pack(src), 0.05, origin=origin, T=T)
    moved = orc.transform(f64(src), T)
    oc, ok, on = orc.voxel_downsample(moved, 0.05, origin=origin)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    np.testing.assert_array_equal(counts(cent), on)
    np.testing.assert_array_equal(xyz(cent), oc.astype(np.float32))


def test_voxel_downsample_edge_cases(engine, b2):
    one = np.array([[0.1, -0.2, 0.3]], dtype=np.float32)
    cent, keys = engine.voxel_downsample(engine.pack(one), 0.02)
    assert cent.shape[0] == 1 and keys.cpu().numpy().tolist() == [[0, 0, 0]]
    np.testing.assert_array_equal(xyz(cent), one)
    same = np.tile(one, (1000, 1))
    cent, keys = engine.voxel_downsample(engine.pack(same), 0.02)
    assert cent.shape[0] == 1 and counts(cent)[0] == 1000
    empty = engine.empty_points(0)
    cent, keys = engine.voxel_downsample(empty, 0.02)
    assert cent.shape[0] == 0
    rng = np.random.default_rng(0)
    neg = rng.uniform(-3, 3, size=(20000, 3)).astype(np.float32)
    cent, keys = engine.voxel_downsample(engine.pack(neg), 0.11, origin=[0.0, 0.0, 0.0])
    oc, ok, on = orc.voxel_downsample(f64(neg), 0.11, origin=[0.0, 0.0, 0.0])
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    assert (ok < 0).any()
    with pytest.raises(b2.B2Error):
        engine.voxel_downsample(engine.pack(neg), 0.0)
    far = np.array([[0, 0, 0], [1e9, 0, 0]], dtype=np.float32)
    with pytest.raises(b2.B2Error, match="out of range"):  # Open3D: "voxel_size is too small"
        engine.voxel_downsample(engine.pack(far), 0.001)
    # the handle stays usable after an error
    cent, _ = engine.voxel_downsample(engine.pack(one), 0.02)
    assert cent.shape[0] == 1


def test_voxel_downsample_wide_key_range_retries(engine):
    """1 mm voxels over metres need > 32 key bits: the radix pass budget is widened transparently
    (icp_processor.py:132-133 uses voxel 0.001 on the whole map)."""
    rng = np.random.default_rng(1)
    pts = rng.uniform(-4, 4, size=(30000, 3)).astype(np.float32)
    cent, keys = engine.voxel_downsample(engine.pack(pts), 0.001)
    oc, ok, on = orc.voxel_downsample(f64(pts), 0.001)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
    np.testing.assert_array_equal(xyz(cent), oc.astype(np.float32))


# ------------------------------------------------------------------ correspondence search (A.3/A.4)
def test_nn_search_bit_exact(engine, pair0, driver):
    src, tgt, _ = pair0
    sd, _ = engine.voxel_downsample(engine.pack(src), 0.02)
    td, _ = engine.voxel_downsample(engine.pack(tgt), 0.02)
    T = np.eye(4)
    T[:3, 3] = [0.011, -0.004, 0.006]
    idx, d2 = engine.nn_search(sd, td, 0.05, T=T)
    oidx, od2 = orc.nn_search(orc.transform(f64(xyz(sd)), T), f64(xyz(td)), 0.05)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_array_equal(d2.cpu().numpy(), od2)
    assert (oidx >= 0).mean() > 0.5 and (oidx < 0).any()


def test_nn_search_raw_scan_target_and_no_transform(engine, pair0, driver):
    src, tgt, _ = pair0
    idx, d2 = engine.nn_search(engine.pack(src[::3].copy()), engine.pack(tgt), 0.05)
    oidx, od2 = orc.nn_search(f64(src[::3]), f64(tgt), 0.05)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_array_equal(d2.cpu().numpy(), od2)


def test_nn_search_ties_strict_radius_and_misses(engine, driver):
    tgt = np.array([[1, 0, 0], [-1, 0, 0], [0, 2, 0], [50, 50, 50]], dtype=np.float32)
    src = np.array([[0, 0, 0], [0, 1.5, 0], [20, 20, 20], [49.5, 50, 50]], dtype=np.float32)
    for radius in (1.5, 1.0, 0.6):
        idx, d2 = engine.nn_search(engine.pack(src), engine.pack(tgt), radius)
        oidx, od2 = orc.nn_search(f64(src), f64(tgt), radius, brute=True)
        np.testing.assert_array_equal(idx.cpu().numpy(), oidx)  # exact tie -> lowest index; d2 < r2 strict
        np.testing.assert_array_equal(d2.cpu().numpy(), od2)


# ------------------------------------------------------------------ ICP (A.3, A.5, A.6)
def _down(engine, a, v=0.02):
    d, _ = engine.voxel_downsample(engine.pack(a), v)
    return d


def test_icp_point_to_point_config1(engine, b2, pair0, driver):
    """BASELINE config 1: two 49k-pt scans, reference parameters, 30 iterations."""
    src, tgt, _ = pair0
    sd, td = _down(engine, src), _down(engine, tgt)
    res, corr = engine.icp(sd, td, b2.make_params(0.05, 30), want_corr=True)
    ores = orc.registration_icp(f64(xyz(sd)), f64(xyz(td)), 0.05, max_iteration=30)
    assert_transform_close(res.matrix(), ores.transformation)
    assert res.iterations == ores.iterations
    assert res.fitness == ores.fitness
    assert res.inlier_rmse == pytest.approx(ores.inlier_rmse, rel=1e-9)
    np.testing.assert_array_equal(corr.cpu().numpy(), ores.correspondence)


def test_icp_with_init_and_convergence_stop(engine, b2, pair0):
    src, tgt, T_true = pair0
    sd, td = _down(engine, src), _down(engine, tgt)
    res, _ = engine.icp(sd, td, b2.make_params(0.05, 500), init=T_true)
    ores = orc.registration_icp(f64(xyz(sd)), f64(xyz(td)), 0.05, init=T_true, max_iteration=500)
    assert res.iterations == ores.iterations < 500
    assert_transform_close(res.matrix(), ores.transformation)
    assert_transform_close(res.matrix(), T_true, rad=1e-2, metres=2e-2)  # and it is near the true motion


def test_icp_iteration_semantics(engine, b2, pair0, driver):
    src, tgt, _ = pair0
    sd, td = _down(engine, src), _down(engine, tgt)
    for it in (0, 1, 3):
        res, _ = engine.icp(sd, td, b2.make_params(0.05, it))
        ores = orc.registration_icp(f64(xyz(sd)), f64(xyz(td)), 0.05, max_iteration=it)
        assert res.iterations == ores.iterations == it
        assert res.fitness == ores.fitness
        np.testing.assert_allclose(res.matrix(), ores.transformation, atol=1e-12)


def test_icp_no_overlap(engine, b2, pair0, driver):
    src, tgt, _ = pair0
    sd, td = _down(engine, src + np.float32(30.0)), _down(engine, tgt)
    res, corr = engine.icp(sd, td, b2.make_params(0.05, 30), want_corr=True)
    assert res.fitness == 0.0 and res.inlier_rmse == 0.0 and res.iterations == 1
    np.testing.assert_array_equal(res.matrix(), np.eye(4))
    assert (corr.cpu().numpy() == -1).all()


def test_icp_point_to_plane_config3(engine, b2, pair0, driver):
    """BASELINE config 3: point-to-plane with on-GPU PCA normals (k = 20)."""
    src, tgt, T_true = pair0
    sd, td = _down(engine, src), _down(engine, tgt)
    nrm = engine.estimate_normals(td, 0.04, 20)
    res, corr = engine.icp(sd, td, b2.make_params(0.05, 30, mode=b2.P2L), tgt_normals=nrm, want_corr=True)
    ores = orc.registration_icp(f64(xyz(sd)), f64(xyz(td)), 0.05, mode="p2l", tgt_normals=f64(xyz(nrm)),
                                max_iteration=30)
    assert res.iterations == ores.iterations
    assert_transform_close(res.matrix(), ores.transformation)
    np.testing.assert_array_equal(corr.cpu().numpy(), ores.correspondence)
    assert_transform_close(res.matrix(), T_true, rad=1e-2, metres=2e-2)
    with pytest.raises(b2.B2Error, match="normals"):
        engine.icp(sd, td, b2.make_params(0.05, 30, mode=b2.P2L))


def test_icp_recovers_known_motion_property(engine, b2):
    """Size-independent property: ICP(T*X, X) ~= T^-1 for a small rigid motion."""
    from lidar_slam_b200 import synth

    scene = synth.make_room_scene(3)
    pose = synth.look_at_pose([1.2, 0.9, 1.3], yaw=0.8, pitch=-0.15)
    X = synth.render_scan(scene, pose, seed=5, noise_sigma=0.0).numpy()
    T = np.eye(4)
    a = 0.004
    T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    T[:3, 3] = [0.008, -0.006, 0.005]
    Y = orc.transform(f64(X), T).astype(np.float32)
    res, _ = engine.icp(_down(engine, Y), engine.pack(X), b2.make_params(0.05, 200))
    assert_transform_close(res.matrix(), np.linalg.inv(T), rad=3e-3, metres=6e-3)


# ------------------------------------------------------------------ normals (A.2)
def _angle_up_to_sign(a, b):
    """Angle between directions, ignoring sign; atan2 form (arccos loses all precision near 0)."""
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), np.abs((a * b).sum(1)))


def test_normals_match_oracle_up_to_sign(engine, pair0):
    _, tgt, _ = pair0
    td = _down(engine, tgt)
    nrm = xyz(engine.estimate_normals(td, 0.04, 20)).astype(np.float64)
    onrm = orc.estimate_normals(f64(xyz(td)), 0.04, 20)
    ang = _angle_up_to_sign(nrm, onrm)
    # 1e-4 rad except where the covariance is near-degenerate (two close eigenvalues make the
    # eigenvector ill-conditioned for ANY solver) or a kNN tie sits at the radius
    assert np.mean(ang < 1e-4) > 0.995, np.mean(ang < 1e-4)
    np.testing.assert_allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-6)


def test_normals_knn_cap_and_fallback(engine):
    g = np.arange(0, 0.6, 0.01, dtype=np.float32)  # 1 cm grid: far more than 20 neighbours inside 4 cm
    xx, yy = np.meshgrid(g, g)
    plane = np.stack([xx.ravel(), yy.ravel(), np.full(xx.size, 0.5, np.float32) + 0.001 * np.sin(40 * xx.ravel())], 1)
    nrm = xyz(engine.estimate_normals(engine.pack(plane.astype(np.float32)), 0.04, 20)).astype(np.float64)
    onrm = orc.estimate_normals(f64(plane), 0.04, 20)
    ang = _angle_up_to_sign(nrm, onrm)
    assert np.mean(ang < 1e-4) > 0.97  # grid points are full of exact-distance ties at the 20th neighbour
    lonely = np.array([[0, 0, 0], [5, 0, 0], [10, 0, 0]], dtype=np.float32)
    np.testing.assert_array_equal(xyz(engine.estimate_normals(engine.pack(lonely), 0.5, 20)), [[0, 0, 1]] * 3)


# ------------------------------------------------------------------ resident map (fusion)
@pytest.mark.parametrize("voxel,max_corr", [(0.05, 0.05), (0.02, 0.05)])
def test_map_fuse_export_matches_oracle(engine, pair0, voxel, max_corr):
    src, tgt, T = pair0
    origin = np.array([-9.0, -9.0, -9.0])
    engine.map_init(voxel, origin, max_corr, 1 << 18)
    engine.map_fuse(engine.pack(tgt))
    engine.map_fuse(engine.pack(src), T)
    cent, keys = engine.map_export()
    allpts = np.concatenate([f64(tgt), orc.transform(f64(src), T)])
    oc, ok, on = orc.voxel_downsample(allpts, voxel, origin=origin)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)           # keys + occupancy bit-exact
    np.testing.assert_array_equal(counts(cent), on)
    np.testing.assert_allclose(xyz(cent), oc, rtol=CENT_RTOL, atol=CENT_ATOL)
    assert np.abs(xyz(cent) - oc).max() < 1e-6                       # in practice: float32 rounding only
    assert engine.map_size() == len(ok)
    engine.map_reset()
    assert engine.map_size() == 0


def test_map_icp_matches_oracle(engine, b2, pair0, driver, dense_driver):
    src, tgt, _ = pair0
    origin = np.array([-9.0, -9.0, -9.0])
    for voxel in (0.05, 0.02):
        engine.map_init(voxel, origin, 0.05, 1 << 18)
        engine.map_fuse(engine.pack(tgt))
        cent, _ = engine.map_export()
        sd = _down(engine, src)
        res = engine.map_icp(sd, b2.make_params(0.05, 30))
        ores = orc.registration_icp(f64(xyz(sd)), f64(xyz(cent)), 0.05, max_iteration=30)
        assert res.iterations == ores.iterations and res.fitness == ores.fitness
        assert_transform_close(res.matrix(), ores.transformation)
    with pytest.raises(b2.B2Error, match="exceeds"):
        engine.map_icp(sd, b2.make_params(0.5, 30))


def _check_dense_correspondences(engine, b2, sd, budget, init=None, rel=1e-6, gate=0.05, driver=None):
    """The correspondence set of the last evaluated transform (the result's) of the dense-map step kernel versus the
    oracle's search with that very transform on the exported centroids: matched / unmatched, d2 and the matched
    voxel (ties -> lowest canonical index) bit for bit."""
    cent, _ = engine.map_export()
    l0 = engine.launch_count()
    res, d2, match = engine.map_icp_corr(sd, b2.make_params(gate, budget, rel_fitness=rel, rel_rmse=rel), init=init)
    if driver is not None and budget >= 6:  # the scheme asked for is the one that ran: 1 launch vs budget + 2
        n_launch = engine.launch_count() - l0
        assert (n_launch <= 5) == (driver == "persistent"), (driver, n_launch)
    tgt = f64(xyz(cent))
    oidx, od2 = orc.nn_search(orc.transform(f64(xyz(sd)), res.matrix()), tgt, gate)
    m = match.cpu().numpy()
    np.testing.assert_array_equal(m[:, 3] > 0, oidx >= 0)
    np.testing.assert_array_equal(d2.cpu().numpy(), od2)
    hit = oidx >= 0
    np.testing.assert_array_equal(m[hit, :3], xyz(cent)[oidx[hit]])
    assert res.n_correspondences == int(hit.sum())
    return res, hit


@pytest.mark.parametrize("budget", [0, 1, 2, 3, 6, 12, 30])
def test_dense_map_correspondences_of_the_result_transform_are_bit_exact(engine, b2, pair0, budget, dense_driver):
    """Every launch of the dense-map ICP step after the first searches from what a speculative pass with the
    PREVIOUS transform kept (b2_icp_dense.cuh); stopping the registration after `budget` steps exposes the search of
    that step, large early updates and small late ones alike."""
    src, tgt, _ = pair0
    engine.map_init(0.05, [-9.0, -9.0, -9.0], 0.05, 1 << 18)
    engine.map_fuse(engine.pack(tgt))
    sd = _down(engine, src)
    res, hit = _check_dense_correspondences(engine, b2, sd, budget, rel=0.0 if budget < 30 else 1e-6, driver=dense_driver)
    assert res.iterations == budget or budget == 30
    assert 0.3 < hit.mean() < 1.0  # matched and unmatched queries both occur
    # a start far from the answer: the first updates move queries across several cells
    T0 = np.eye(4)
    T0[:3, 3] = [0.06, -0.05, 0.04]
    _check_dense_correspondences(engine, b2, sd, budget, init=T0, rel=0.0)


def test_dense_map_correspondences_with_more_queries_than_one_round(engine, b2, pair0, dense_driver):
    """A raw 49,152-pt source is more than the 148 x 190 queries the step kernel serves per round: the first round
    searches from the speculative pass, the second one takes the full path — both must agree with the oracle.  Also
    tiny sources (fewer queries than CTAs, than lanes of a warp)."""
    src, tgt, _ = pair0
    engine.map_init(0.05, [-9.0, -9.0, -9.0], 0.05, 1 << 18)
    engine.map_fuse(engine.pack(tgt))
    for budget in (1, 4):
        _check_dense_correspondences(engine, b2, engine.pack(src), budget, rel=0.0)
    for n in (1, 2, 11, 149, 1481):
        _check_dense_correspondences(engine, b2, engine.pack(src[:: len(src) // n][:n].copy()), 3, rel=0.0)


@pytest.mark.parametrize("seed", [0, 1])
def test_dense_map_correspondences_on_lattice_clouds(engine, b2, seed, dense_driver):
    """Voxel centroids on a 1/16 m lattice and queries on the half-lattice (every coordinate exactly representable):
    distances tie exactly between two, four and eight voxels, queries sit exactly on voxel faces and neighbours lie
    at exactly the gate (rejected: d2 < r2 is strict) — under the identity, a lattice-preserving start and a generic
    one."""
    rng = np.random.default_rng(seed)
    v = 0.0625
    tgt = ((rng.integers(-30, 30, (20000, 3)) + 0.5) * v).astype(np.float32)  # voxel centres, many voxels hit twice
    src = (rng.integers(-60, 60, (6000, 3)) * (v / 2)).astype(np.float32)
    engine.map_init(v, [-4.0, -4.0, -4.0], v, 1 << 18)
    engine.map_fuse(engine.pack(tgt))
    sd = engine.pack(src)
    for budget in (0, 1, 4):
        _, hit = _check_dense_correspondences(engine, b2, sd, budget, rel=0.0, gate=v)
        assert hit.any() and (~hit).any()
    T0 = np.eye(4)
    T0[:3, 3] = [v / 4, 0.0, -v / 2]
    _check_dense_correspondences(engine, b2, sd, 2, init=T0, rel=0.0, gate=v)
    T0[:3, 3] = [0.0131, -0.0072, 0.0205]
    _check_dense_correspondences(engine, b2, sd, 3, init=T0, rel=0.0, gate=v)


def test_map_capacity_error(engine, b2, pair0):
    src, _, _ = pair0
    engine.map_init(0.02, [-9.0, -9.0, -9.0], 0.02, 1024)
    with pytest.raises(b2.B2Error, match="full"):
        engine.map_fuse(engine.pack(src))
    engine.map_init(0.05, [-9.0, -9.0, -9.0], 0.05, 1 << 16)  # re-initialising recovers


def test_slam_sequence_resident_mode(engine, b2, dense_driver):
    """Streaming scan->map ICP + fuse (BASELINE config 2/4 shape) against the oracle's resident twin."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(1, length=12.0)
    poses = synth.corridor_trajectory(scene, 6, seed=1)
    origin = np.array([-20.0, -20.0, -20.0])
    engine.map_init(0.05, origin, 0.05, 1 << 18)
    ref = pipeline.ResidentSequence(0.05, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30)
    prm = b2.make_params(0.05, 30)
    for k, pose in enumerate(poses):
        scan = synth.render_scan(scene, pose, seed=100 + k).numpy()
        res = engine.slam_scan(scan, 0.02, prm)
        ores = ref.process(scan)
        if k == 0:
            continue
        assert res.iterations == ores.iterations, k
        assert_transform_close(res.matrix(), ores.transformation)
        assert res.fitness == pytest.approx(ores.fitness, abs=2e-4)
    cent, keys = engine.map_export()
    oc, ok = ref.map_points()
    assert len(ok) > 3000
    # after several chained registrations a point within float rounding of a voxel face may land in
    # the neighbouring voxel: count and bound those instead of hiding them
    gk = {tuple(k) for k in keys.cpu().numpy().tolist()}
    rk = {tuple(k) for k in ok.tolist()}
    assert len(gk ^ rk) <= max(4, len(rk) // 2000)
    np.testing.assert_allclose(engine.get_pose(), ref.pose, atol=1e-6)


def test_batch_icp_matches_per_pair_oracle(engine, b2, driver):
    from lidar_slam_b200 import synth

    pairs = [synth.make_pair(s) for s in (0, 1, 2)]
    srcs = [p[0].numpy()[::2].copy() for p in pairs]
    tgts = [p[1].numpy()[::2].copy() for p in pairs]
    so = np.cumsum([0] + [len(s) for s in srcs]).astype(np.int32)
    to = np.cumsum([0] + [len(t) for t in tgts]).astype(np.int32)
    S = engine.pack(np.concatenate(srcs))
    Tt = engine.pack(np.concatenate(tgts))
    res = engine.icp_batch(S, torch.from_numpy(so).cuda(), Tt, torch.from_numpy(to).cuda(),
                           b2.make_params(0.05, 30), ds_voxel=0.02)
    engine.synchronize()
    rec = b2.results_to_numpy(res)
    for i in range(len(pairs)):
        sd, _, _ = orc.voxel_downsample(f64(srcs[i]), 0.02)
        td, _, _ = orc.voxel_downsample(f64(tgts[i]), 0.02)
        ores = orc.registration_icp(f64(sd), f64(td), 0.05, max_iteration=30)
        assert rec["iterations"][i] == ores.iterations
        assert rec["fitness"][i] == ores.fitness
        assert_transform_close(rec["transformation"][i], ores.transformation)


def test_transform_and_pack_roundtrip(engine, pair0):
    src, _, T = pair0
    p = engine.pack(src)
    np.testing.assert_array_equal(engine.unpack(p), src)
    moved = engine.transform(p, T)
    engine.synchronize()
    np.testing.assert_array_equal(xyz(moved), orc.transform(f64(src), T).astype(np.float32))


def test_projection_kernel_matches_host(engine):
    from lidar_slam_b200 import synth

    rng = np.random.default_rng(3)
    rec = np.stack([rng.integers(0, 256, 4000), rng.integers(0, 192, 4000), rng.uniform(0.3, 5, 4000)], 1).astype(np.float32)
    fx, fy, cx, cy = synth.intrinsics()
    for numpy2, mode in ((False, "f64"), (True, "f32")):
        got = xyz(engine.project_pixels(rec, fx, fy, cx, cy, numpy2=numpy2))
        np.testing.assert_array_equal(got, synth.project_records(torch.from_numpy(rec), mode=mode).numpy())


# ------------------------------------------------------------------ map maintenance filters (A.8)
def _cloud_with_outliers(seed=0, n=24000):
    rng = np.random.default_rng(seed)
    g = rng.uniform(0, 1, size=(n, 2))
    wall = np.stack([2.0 * g[:, 0], rng.normal(scale=0.003, size=n), 1.5 * g[:, 1]], 1)
    floor = np.stack([2.0 * g[:, 1], 1.2 * g[:, 0], rng.normal(scale=0.003, size=n)], 1)
    stray = rng.uniform(-0.5, 2.5, size=(300, 3))
    far = np.array([[9.0, 9.0, 9.0], [9.02, 9.0, 9.0], [-7.0, 3.0, 1.0]])
    return np.concatenate([wall, floor, stray, far]).astype(np.float32)


def test_radius_outlier_filter_matches_oracle(engine):
    pts = _cloud_with_outliers(1)
    keep, kept = engine.radius_outlier_mask(engine.pack(pts), 8, 0.023)
    omask = orc.radius_outlier_mask(f64(pts), 8, 0.023)
    np.testing.assert_array_equal(keep.cpu().numpy().astype(bool), omask)
    assert kept == int(omask.sum()) and 0 < kept < len(pts)


@pytest.mark.parametrize("k,ratio", [(45, 3.5), (80, 2.0)])
def test_statistical_outlier_filter_matches_oracle(engine, k, ratio):
    pts = _cloud_with_outliers(2)
    keep, kept = engine.statistical_outlier_mask(engine.pack(pts), k, ratio)
    omask = orc.statistical_outlier_mask(f64(pts), k, ratio)
    g = keep.cpu().numpy().astype(bool)
    # identical except for points whose mean distance sits within rounding of the threshold
    assert int((g != omask).sum()) <= 1, int((g != omask).sum())
    assert abs(kept - int(omask.sum())) <= 1 and 0 < kept < len(pts)
    # explicit (deliberately poor) cell sizes exercise the ring growth and the overflow path
    for cell in (0.01, 0.2):
        keep2, _ = engine.statistical_outlier_mask(engine.pack(pts[::4].copy()), k, ratio, cell=cell)
        o2 = orc.statistical_outlier_mask(f64(pts[::4]), k, ratio)
        assert int((keep2.cpu().numpy().astype(bool) != o2).sum()) <= 1


def test_statistical_filter_small_cloud(engine):
    pts = np.array([[0, 0, 0], [0.01, 0, 0], [0, 0.01, 0], [0.5, 0.5, 0.5]], dtype=np.float32)
    keep, kept = engine.statistical_outlier_mask(engine.pack(pts), 45, 1.0)  # k > n: every point is a neighbour
    np.testing.assert_array_equal(keep.cpu().numpy().astype(bool), orc.statistical_outlier_mask(f64(pts), 45, 1.0))


def test_select_points_preserves_order(engine):
    pts = _cloud_with_outliers(3)
    keep, kept = engine.radius_outlier_mask(engine.pack(pts), 8, 0.023)
    sel = engine.select_points(engine.pack(pts), keep)
    engine.synchronize()
    np.testing.assert_array_equal(xyz(sel), pts[keep.cpu().numpy().astype(bool)])
    assert sel.shape[0] == kept


def test_batch_icp_ragged_and_empty_pairs(engine, b2, driver):
    """Ragged batch: a full pair, a tiny pair, a pair with an EMPTY source and a non-overlapping pair."""
    from lidar_slam_b200 import synth

    s0, t0, _ = synth.make_pair(5)
    s0, t0 = s0.numpy()[::3].copy(), t0.numpy()[::3].copy()
    srcs = [s0, s0[:50].copy(), np.zeros((0, 3), np.float32), (s0 + np.float32(40.0))]
    tgts = [t0, t0, t0[:2000].copy(), t0]
    so = np.cumsum([0] + [len(s) for s in srcs]).astype(np.int32)
    to = np.cumsum([0] + [len(t) for t in tgts]).astype(np.int32)
    res = engine.icp_batch(engine.pack(np.concatenate(srcs)), torch.from_numpy(so).cuda(),
                           engine.pack(np.concatenate(tgts)), torch.from_numpy(to).cuda(), b2.make_params(0.05, 20))
    engine.synchronize()
    rec = b2.results_to_numpy(res)
    for i in (0, 1, 3):
        ores = orc.registration_icp(f64(srcs[i]), f64(tgts[i]), 0.05, max_iteration=20)
        assert rec["iterations"][i] == ores.iterations, i
        assert rec["fitness"][i] == ores.fitness, i
        assert_transform_close(rec["transformation"][i], ores.transformation)
    assert rec["fitness"][2] == 0.0 and rec["n_correspondences"][2] == 0
    np.testing.assert_array_equal(rec["transformation"][2], np.eye(4))
    assert rec["fitness"][3] == 0.0


def test_slam_sequence_with_varying_scan_sizes(engine, b2):
    """Real scans keep ~half of the pixels (ARDepthViewController.swift:112), a different count every frame:
    the graph-replayed step reads the live count from device memory and must still match the oracle."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(21, length=12.0)
    poses = synth.corridor_trajectory(scene, 7, seed=5)
    origin = np.array([-20.0, -20.0, -20.0])
    engine.map_init(0.05, origin, 0.05, 1 << 18)
    ref = pipeline.ResidentSequence(0.05, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=25)
    prm = b2.make_params(0.05, 25)
    sizes = []
    for k, pose in enumerate(poses):
        scan = synth.render_scan(scene, pose, seed=300 + k, keep_prob=0.5).numpy()
        sizes.append(len(scan))
        res = engine.slam_scan(scan, 0.02, prm)
        ores = ref.process(scan)
        if k == 0:
            continue
        assert res.iterations == ores.iterations, k
        assert res.n_correspondences == int((ores.correspondence >= 0).sum())
        assert_transform_close(res.matrix(), ores.transformation)
    assert len(set(sizes)) == len(sizes) and max(sizes) < 30000


def test_deferred_results_equal_per_scan_results(engine, b2):
    """b2_slam_scan with out == NULL + b2_slam_results (uploads double-buffered on the copy stream) gives, scan
    for scan, exactly what the synchronous call returns."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(11, length=10.0)
    poses = synth.corridor_trajectory(scene, 7, seed=11, step=0.02)
    scans = [np.ascontiguousarray(synth.render_scan(scene, p, seed=70 + k).numpy()) for k, p in enumerate(poses)]
    prm = b2.make_params(0.05, 30)
    origin = np.array([-20.0, -20.0, -20.0])
    engine.map_init(0.05, origin, 0.05, 1 << 18)
    sync = [engine.slam_scan(s, 0.02, prm) for s in scans]
    map_a, _ = engine.map_export()
    map_a = map_a.clone()
    engine.map_init(0.05, origin, 0.05, 1 << 18)
    engine.set_pose(np.eye(4))
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]  # truly asynchronous uploads
    for p in pinned:
        assert engine.slam_scan(p.numpy(), 0.02, prm, want_result=False) is None
    got = engine.slam_results(len(scans))
    for a, b in zip(sync, got):
        np.testing.assert_array_equal(a.matrix(), b.matrix())
        assert (a.fitness, a.inlier_rmse, a.iterations, a.n_correspondences) == \
               (b.fitness, b.inlier_rmse, b.iterations, b.n_correspondences)
    assert got[0].iterations == 0 and got[0].fitness == 0.0  # the first scan becomes the map
    map_b, _ = engine.map_export()
    assert torch.equal(map_a, map_b)
    assert [r.iterations for r in engine.slam_results(3)] == [r.iterations for r in sync[-3:]]
    with pytest.raises(b2.B2Error, match="results asked"):
        engine.slam_results(100000)


@pytest.mark.parametrize("budget,rel", [(500, 1e-6), (57, 0.0)])
def test_slam_sequence_long_iteration_budget_uses_loop_node(engine, b2, budget, rel, dense_driver):
    """Long iteration budgets under both launch schemes: the persistent launch simply loops longer; with chained
    launches the budget above the 40-launch window runs as a CUDA-graph WHILE node once the step is captured (third
    scan on).  The reference default of 500 iterations must stop at convergence exactly like the oracle, and with
    the convergence test disabled (rel = 0) every one of the 57 steps must run (40 chained + 17 loop-node trips)."""
    from lidar_slam_b200 import synth

    scene = synth.make_corridor_scene(13, length=12.0)
    poses = synth.corridor_trajectory(scene, 6, seed=13)
    origin = np.array([-20.0, -20.0, -20.0])
    engine.map_init(0.05, origin, 0.05, 1 << 18)
    engine.set_pose(np.eye(4))
    ref = pipeline.ResidentSequence(0.05, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=budget,
                                    relative_fitness=rel, relative_rmse=rel)
    prm = b2.make_params(0.05, budget, rel_fitness=rel, rel_rmse=rel)
    for k, pose in enumerate(poses):
        scan = synth.render_scan(scene, pose, seed=300 + k).numpy()
        res = engine.slam_scan(scan, 0.02, prm)
        ores = ref.process(scan)
        if k == 0:
            continue
        assert res.iterations == ores.iterations, k
        if rel == 0.0:
            assert res.iterations == budget
        assert_transform_close(res.matrix(), ores.transformation)
        assert res.fitness == pytest.approx(ores.fitness, abs=2e-4)


def _one_step(engine, b2, src, tgt, gate, driver_name="group"):
    engine.set_icp_driver(driver_name)
    try:
        res, corr = engine.icp(engine.pack(src.astype(np.float32)), engine.pack(tgt.astype(np.float32)),
                               b2.make_params(gate, 1), want_corr=True)
    finally:
        engine.set_icp_driver("auto")
    ores = orc.registration_icp(f64(src.astype(np.float32)), f64(tgt.astype(np.float32)), gate, max_iteration=1)
    np.testing.assert_array_equal(corr.cpu().numpy(), ores.correspondence)
    return res, ores


@pytest.mark.parametrize("case", ["coplanar", "reflection", "three_points", "coplanar_sweep"])
def test_se3_solve_special_covariances(engine, b2, case):
    """The on-device Umeyama step away from the fast path (Newton polar iteration needs det(C) > 0): rank-2
    covariances (coplanar clouds, three points) and a covariance whose best orthogonal fit is a reflection
    must come out as the proper rotation Eigen's umeyama / the oracle returns."""
    rng = np.random.default_rng(42)
    if case.startswith("coplanar"):
        tgt = np.concatenate([rng.uniform(0.0, 4.0, (400, 2)), np.full((400, 1), 1.25)], axis=1)
    elif case == "three_points":
        tgt = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.2], [0.1, 1.3, -0.1]])
    else:
        tgt = rng.uniform(-2.0, 2.0, (300, 3))
    a = 0.02
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    src = (tgt - [0.01, -0.02, 0.0]) @ R  # small in-plane motion: nearest neighbours are the true pairs
    gate = 0.5
    if case == "reflection":
        # squash and mirror the source through the z = 0 plane with noise: the unconstrained optimum has det < 0
        tgt = tgt * [1.0, 1.0, 0.02]
        src = tgt * [1.0, 1.0, -1.0] + rng.normal(scale=1e-3, size=tgt.shape)
    res, ores = _one_step(engine, b2, src, tgt, gate, "sweep" if case.endswith("sweep") else "group")
    assert res.fitness == ores.fitness == 1.0
    T = res.matrix()
    assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-9
    np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-9)
    assert_transform_close(T, ores.transformation)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_lattice_clouds_exact_boundaries_and_ties(engine, seed, driver):
    """Points on a 1/8 m lattice (every coordinate exactly representable): voxel faces are hit exactly, many
    points coincide, and nearest-neighbour distances tie exactly — keys, occupancy, correspondences and d2 must
    still be bit-identical to the oracle (ties -> lowest target index; d2 < r2 strict)."""
    rng = np.random.default_rng(seed)
    tgt = (rng.integers(-40, 40, (6000, 3)) * 0.125).astype(np.float32)
    src = (rng.integers(-40, 40, (3000, 3)) * 0.125).astype(np.float32)
    for voxel in (0.25, 0.125, 0.3):
        cent, keys = engine.voxel_downsample(engine.pack(tgt), voxel)
        oc, ok, on = orc.voxel_downsample(f64(tgt), voxel)
        np.testing.assert_array_equal(keys.cpu().numpy(), ok)
        np.testing.assert_array_equal(counts(cent), on)
        np.testing.assert_allclose(xyz(cent), oc, rtol=CENT_RTOL, atol=CENT_ATOL)
    for radius in (0.125, 0.25, 0.3):  # 0.125 / 0.25: neighbours at exactly the radius must be rejected
        idx, d2 = engine.nn_search(engine.pack(src), engine.pack(tgt), radius)
        oidx, od2 = orc.nn_search(f64(src), f64(tgt), radius)
        np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
        np.testing.assert_array_equal(d2.cpu().numpy(), od2)
        assert (oidx >= 0).any() and (oidx < 0).any()


def test_batch_normals_and_point_to_plane_match_per_pair_oracle(engine, b2, driver):
    """b2_estimate_normals_batch / b2_icp_batch_p2l (config 3 over a batch): normals searched inside each cloud
    only; every pair's registration equals the single-pair oracle run with that pair's normals."""
    from lidar_slam_b200 import synth

    pairs = [synth.make_pair(s) for s in (3, 4, 5)]
    srcs = [p[0].numpy()[::2].copy() for p in pairs]
    tgts = [p[1].numpy()[::2].copy() for p in pairs]
    tgts[1] = tgts[1][:9000].copy()  # ragged
    so = np.cumsum([0] + [len(s) for s in srcs]).astype(np.int32)
    to = np.cumsum([0] + [len(t) for t in tgts]).astype(np.int32)
    T4 = engine.pack(np.concatenate(tgts))
    # batched normals on the raw targets == per-cloud normals
    nb = engine.estimate_normals_batch(T4, torch.from_numpy(to).cuda(), 0.04, 20)
    for i, t in enumerate(tgts):
        ni = engine.estimate_normals(engine.pack(t), 0.04, 20)
        np.testing.assert_array_equal(xyz(nb)[to[i]:to[i + 1]], xyz(ni))
    # batched point-to-plane with internal down-sampling and normals
    prm = b2.make_params(0.05, 30, mode=b2.P2L)
    res = engine.icp_batch(engine.pack(np.concatenate(srcs)), torch.from_numpy(so).cuda(), T4, torch.from_numpy(to).cuda(),
                           prm, ds_voxel=0.02, normal_radius=0.04, normal_max_nn=20)
    engine.synchronize()
    rec = b2.results_to_numpy(res)
    for i in range(len(pairs)):
        sd, _, _ = orc.voxel_downsample(f64(srcs[i]), 0.02)
        td, _, _ = orc.voxel_downsample(f64(tgts[i]), 0.02)
        td32 = td.astype(np.float32)
        nrm = engine.estimate_normals(engine.pack(td32), 0.04, 20)  # (GPU normals == oracle up to sign: tested above)
        ores = orc.registration_icp(f64(sd.astype(np.float32)), f64(td32), 0.05, mode="p2l", tgt_normals=f64(xyz(nrm)),
                                    max_iteration=30)
        assert rec["iterations"][i] == ores.iterations, i
        assert rec["fitness"][i] == ores.fitness, i
        assert_transform_close(rec["transformation"][i], ores.transformation)
    with pytest.raises(b2.B2Error, match="normal_radius"):
        engine.icp_batch(engine.pack(np.concatenate(srcs)), torch.from_numpy(so).cuda(), T4,
                         torch.from_numpy(to).cuda(), prm, ds_voxel=0.02)


def test_non_finite_points_fail_loudly_and_do_not_poison_the_engine(engine, b2, pair0):
    """A NaN / Inf coordinate has no voxel (Open3D's int32 voxel index is undefined for it): the call must
    raise, not hang or return garbage, and the engine must stay usable."""
    src, tgt, _ = pair0
    for bad in (np.nan, np.inf, -np.inf, 3e38):
        pts = src[:5000].copy()
        pts[1234, 1] = bad
        with pytest.raises(b2.B2Error):
            engine.voxel_downsample(engine.pack(pts), 0.02)
        with pytest.raises(b2.B2Error):  # a target cannot be gridded either
            engine.icp(engine.pack(src[:5000].copy()), engine.pack(pts), b2.make_params(0.05, 3))
        # a non-finite SOURCE point simply has no correspondence
        idx, _ = engine.nn_search(engine.pack(pts), engine.pack(tgt), 0.05)
        oidx, _ = orc.nn_search(f64(np.delete(pts, 1234, axis=0)), f64(tgt), 0.05)
        got = idx.cpu().numpy()
        assert got[1234] == -1
        np.testing.assert_array_equal(np.delete(got, 1234), oidx)
    cent, keys = engine.voxel_downsample(engine.pack(src), 0.02)  # still fine afterwards
    oc, ok, on = orc.voxel_downsample(f64(src), 0.02)
    np.testing.assert_array_equal(keys.cpu().numpy(), ok)
