"""Known-answer tests of the CPU oracle (oracle/oracle.cpp).

The reference pins no numeric result for this path and Open3D is not
available, so the oracle is pinned by analytically solvable cases authored here and
by an independently written NumPy/SciPy twin (oracle/twin.py).
"""
import numpy as np
import pytest

from oracle import cpu, twin


def rot(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def rigid(axis, ang, t):
    T = np.eye(4)
    T[:3, :3] = rot(axis, ang)
    T[:3, 3] = t
    return T


# ---------------------------------------------------------------- A.1 voxels
def test_voxel_keys_hand_computed():
    # origin defaults to min_bound - v/2 = (-0.5, -0.5, -0.5) for v = 1
    pts = np.array([[0.0, 0.0, 0.0], [0.4, 0.4, 0.4], [0.6, 0.0, 0.0], [2.0, 3.0, 1.0], [0.5, 0.0, 0.0]])
    keys = cpu.voxel_keys(pts, 1.0)
    assert keys.tolist() == [[0, 0, 0], [0, 0, 0], [1, 0, 0], [2, 3, 1], [1, 0, 0]]
    cent, k, cnt = cpu.voxel_downsample(pts, 1.0)
    assert k.tolist() == [[0, 0, 0], [1, 0, 0], [2, 3, 1]]  # canonical (lexicographic) order
    assert cnt.tolist() == [2, 2, 1]
    np.testing.assert_array_equal(cent[0], [0.2, 0.2, 0.2])
    np.testing.assert_array_equal(cent[1], [(0.6 + 0.5) / 2, 0.0, 0.0])
    np.testing.assert_array_equal(cent[2], [2.0, 3.0, 1.0])


def test_voxel_explicit_origin_and_negative_coordinates():
    pts = np.array([[-0.01, 0.0, 0.0], [0.0, 0.0, 0.0], [-1.0, -1.0, -1.0]])
    keys = cpu.voxel_keys(pts, 0.5, origin=[0.0, 0.0, 0.0])
    assert keys.tolist() == [[-1, 0, 0], [0, 0, 0], [-2, -2, -2]]


def test_voxel_downsample_idempotent_at_same_origin():
    rng = np.random.default_rng(1)
    pts = rng.uniform(-2, 2, size=(5000, 3))
    org = np.array([-3.0, -3.0, -3.0])
    c1, k1, _ = cpu.voxel_downsample(pts, 0.25, origin=org)
    c2, k2, n2 = cpu.voxel_downsample(c1, 0.25, origin=org)
    np.testing.assert_array_equal(k1, k2)
    np.testing.assert_array_equal(c1, c2)
    assert (n2 == 1).all()


def test_voxel_permutation_invariant_occupancy():
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(3000, 3))
    org = [-10.0, -10.0, -10.0]
    _, k1, n1 = cpu.voxel_downsample(pts, 0.3, origin=org)
    _, k2, n2 = cpu.voxel_downsample(pts[rng.permutation(len(pts))], 0.3, origin=org)
    np.testing.assert_array_equal(k1, k2)
    np.testing.assert_array_equal(n1, n2)


def test_voxel_matches_twin():
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 3, size=(20000, 3)).astype(np.float32).astype(np.float64)
    c, k, n = cpu.voxel_downsample(pts, 0.05)
    c2, k2, n2 = twin.voxel_downsample(pts, 0.05)
    np.testing.assert_array_equal(k, k2)
    np.testing.assert_array_equal(n, n2)
    np.testing.assert_allclose(c, c2, rtol=0, atol=1e-12)


# ---------------------------------------------------------------- A.4 search
def test_nn_kdtree_equals_brute_force():
    rng = np.random.default_rng(4)
    tgt = rng.uniform(0, 1, size=(4000, 3))
    src = rng.uniform(-0.05, 1.05, size=(3000, 3))
    i1, d1 = cpu.nn_search(src, tgt, 0.05)
    i2, d2 = cpu.nn_search(src, tgt, 0.05, brute=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    assert (i1 < 0).any() and (i1 >= 0).any()


def test_nn_strict_radius_and_tie_break():
    tgt = np.array([[1.0, 0.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 2.0, 0.0]])
    src = np.array([[0.0, 0.0, 0.0]])
    idx, d2 = cpu.nn_search(src, tgt, 1.5)
    assert idx[0] == 0 and d2[0] == 1.0  # exact tie between 0 and 1 -> lowest index
    idx, _ = cpu.nn_search(src, tgt, 1.0)  # d2 < r2 is strict
    assert idx[0] == -1
    idx, _ = cpu.nn_search(src, tgt, 1.0, brute=True)
    assert idx[0] == -1


def test_knn_hybrid_prefix():
    tgt = np.array([[0.0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [0.3, 0, 0], [5.0, 0, 0]])
    idx, d2, cnt = cpu.knn(tgt[:1], tgt, 4, radius=0.25)
    assert cnt[0] == 3 and idx[0, :3].tolist() == [0, 1, 2]
    idx, d2, cnt = cpu.knn(tgt[:1], tgt, 3)
    assert cnt[0] == 3 and idx[0].tolist() == [0, 1, 2]


# ---------------------------------------------------------------- A.5 / A.6 estimation
def test_svd3_against_lapack():
    rng = np.random.default_rng(5)
    for _ in range(200):
        M = rng.normal(size=(3, 3)) * rng.choice([1e-6, 1.0, 1e3])
        U, S, V = cpu.svd3(M)
        np.testing.assert_allclose(U @ np.diag(S) @ V.T, M, atol=1e-12 * max(1.0, np.abs(M).max()))
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(S, np.linalg.svd(M, compute_uv=False), rtol=1e-10, atol=1e-300)
        assert S[0] >= S[1] >= S[2] >= 0


def test_umeyama_recovers_exact_rigid_motion():
    rng = np.random.default_rng(6)
    P = rng.uniform(-1, 1, size=(500, 3))
    T = rigid([1, 2, 3], 0.3, [0.1, -0.2, 0.05])
    Q = P @ T[:3, :3].T + T[:3, 3]
    idx = np.arange(len(P), dtype=np.int32)
    Te = cpu.umeyama(P, Q, idx, idx)
    np.testing.assert_allclose(Te, T, atol=1e-13)
    np.testing.assert_allclose(Te, twin.umeyama(P, Q), atol=1e-13)


def test_umeyama_reflection_case_returns_proper_rotation():
    rng = np.random.default_rng(7)
    P = rng.uniform(-1, 1, size=(200, 3))
    Q = P * np.array([1.0, 1.0, -1.0])  # mirrored: best proper rotation has det +1
    idx = np.arange(len(P), dtype=np.int32)
    Te = cpu.umeyama(P, Q, idx, idx)
    assert np.linalg.det(Te[:3, :3]) == pytest.approx(1.0, abs=1e-12)
    np.testing.assert_allclose(Te, twin.umeyama(P, Q), atol=1e-12)


def test_umeyama_coplanar_points():
    rng = np.random.default_rng(8)
    P = np.concatenate([rng.uniform(-1, 1, size=(300, 2)), np.zeros((300, 1))], axis=1)
    T = rigid([0, 0, 1], 0.2, [0.3, 0.1, 0.0])
    Q = P @ T[:3, :3].T + T[:3, 3]
    idx = np.arange(len(P), dtype=np.int32)
    np.testing.assert_allclose(cpu.umeyama(P, Q, idx, idx), T, atol=1e-12)


def test_empty_correspondences_give_identity():
    P = np.zeros((3, 3))
    e = np.zeros(0, dtype=np.int32)
    np.testing.assert_array_equal(cpu.umeyama(P, P, e, e), np.eye(4))


def test_point_to_plane_step_matches_twin_and_small_motion():
    rng = np.random.default_rng(9)
    Q = rng.uniform(-1, 1, size=(800, 3))
    N = rng.normal(size=(800, 3))
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    T = rigid([0.2, -1, 0.4], 0.01, [0.004, -0.003, 0.002])
    P = (Q - T[:3, 3]) @ T[:3, :3]  # P = T^-1 Q, so the step should recover ~T
    idx = np.arange(len(P), dtype=np.int32)
    Te = cpu.point_to_plane_step(P, Q, N, idx, idx)
    np.testing.assert_allclose(Te, twin.point_to_plane_step(P, Q, N), atol=1e-12)
    np.testing.assert_allclose(Te, T, atol=2e-4)  # linearised step: second-order residual


# ---------------------------------------------------------------- A.2 normals
def test_eig3_smallest_against_lapack():
    rng = np.random.default_rng(10)
    for _ in range(300):
        A = rng.normal(size=(3, 3))
        Cm = A @ np.diag(rng.uniform(0.01, 1.0, 3) * [1e-3, 1.0, 1.0]) @ A.T
        v = cpu.eig3_smallest(Cm)
        w, V = np.linalg.eigh(Cm)
        assert abs(abs(v @ V[:, 0]) - 1.0) < 1e-7, (v, V[:, 0])


def test_eig3_degenerate_inputs():
    np.testing.assert_array_equal(cpu.eig3_smallest(np.eye(3)), [0, 0, 1])
    np.testing.assert_array_equal(cpu.eig3_smallest(np.zeros((3, 3))), [0, 0, 0])
    np.testing.assert_array_equal(cpu.eig3_smallest(np.diag([0.5, 2.0, 3.0])), [1, 0, 0])


def test_normals_of_axis_aligned_plane():
    g = np.arange(0, 1, 0.02)
    xx, yy = np.meshgrid(g, g)
    pts = np.stack([xx.ravel(), yy.ravel(), np.full(xx.size, 0.7)], axis=1)
    n = cpu.estimate_normals(pts, 0.05, 20)
    np.testing.assert_allclose(np.abs(n[:, 2]), 1.0, atol=1e-9)
    np.testing.assert_allclose(n[:, :2], 0.0, atol=1e-6)


def test_normals_isolated_points_fall_back_to_z():
    pts = np.array([[0.0, 0, 0], [10.0, 0, 0], [20.0, 0, 0]])
    np.testing.assert_array_equal(cpu.estimate_normals(pts, 0.5, 20), [[0, 0, 1]] * 3)


def test_normals_match_twin_up_to_sign(pair0):
    src, _, _ = pair0
    d, _, _ = cpu.voxel_downsample(src.astype(np.float64), 0.02)
    d = d[:6000]
    a = cpu.estimate_normals(d, 0.04, 20)
    b = twin.estimate_normals(d, 0.04, 20)
    cosang = np.abs((a * b).sum(1))
    assert np.mean(cosang > 1 - 1e-8) > 0.995  # the rest: near-degenerate covariances / kNN ties


# ---------------------------------------------------------------- A.3 ICP loop
def grid_cloud():
    g = np.arange(0, 1.0001, 0.04)
    xx, yy = np.meshgrid(g, g)
    a = np.stack([xx.ravel(), yy.ravel(), np.zeros(xx.size)], axis=1)
    b = np.stack([xx.ravel(), np.zeros(xx.size), yy.ravel()], axis=1)
    c = np.stack([np.zeros(xx.size), xx.ravel(), yy.ravel()], axis=1)
    return np.concatenate([a, b, c])


def test_icp_recovers_translation_exactly():
    tgt = grid_cloud()
    t = np.array([0.006, -0.004, 0.005])
    src = tgt - t
    r = cpu.registration_icp(src, tgt, 0.02, max_iteration=50)
    np.testing.assert_allclose(r.transformation[:3, 3], t, atol=1e-12)
    np.testing.assert_allclose(r.transformation[:3, :3], np.eye(3), atol=1e-12)
    assert r.fitness == 1.0 and r.inlier_rmse < 1e-12
    assert r.iterations == 2  # one step lands on the solution, the second detects convergence


def test_icp_iteration_semantics():
    tgt = grid_cloud()
    src = tgt - [0.006, -0.004, 0.005]
    r0 = cpu.registration_icp(src, tgt, 0.02, max_iteration=0)
    np.testing.assert_array_equal(r0.transformation, np.eye(4))
    assert r0.iterations == 0 and r0.fitness == 1.0
    r1 = cpu.registration_icp(src, tgt, 0.02, max_iteration=1)
    assert r1.iterations == 1
    init = rigid([0, 0, 1], 0.0, [0.001, 0, 0])
    ri = cpu.registration_icp(src, tgt, 0.02, init=init, max_iteration=0)
    np.testing.assert_array_equal(ri.transformation, init)


def test_icp_no_overlap_returns_init():
    tgt = grid_cloud()
    src = tgt + 10.0
    r = cpu.registration_icp(src, tgt, 0.02, max_iteration=30)
    np.testing.assert_array_equal(r.transformation, np.eye(4))
    assert r.fitness == 0.0 and r.inlier_rmse == 0.0 and r.iterations == 1
    assert (r.correspondence == -1).all()


def test_icp_matches_twin_on_synthetic_pair(pair0):
    src, tgt, _ = pair0
    sd, _, _ = cpu.voxel_downsample(src.astype(np.float64), 0.02)
    td, _, _ = cpu.voxel_downsample(tgt.astype(np.float64), 0.02)
    r = cpu.registration_icp(sd, td, 0.05, max_iteration=10)
    T2, f2, r2, it2, idx2 = twin.registration_icp(sd, td, 0.05, max_iteration=10)
    np.testing.assert_allclose(r.transformation, T2, atol=1e-11)
    assert r.iterations == it2 and r.fitness == f2
    assert r.inlier_rmse == pytest.approx(r2, rel=1e-10)
    assert np.mean(r.correspondence == idx2) == 1.0


def test_icp_p2l_requires_normals_and_matches_twin(pair0):
    src, tgt, _ = pair0
    sd, _, _ = cpu.voxel_downsample(src.astype(np.float64), 0.02)
    td, _, _ = cpu.voxel_downsample(tgt.astype(np.float64), 0.02)
    with pytest.raises(RuntimeError):
        cpu.registration_icp(sd, td, 0.05, mode="p2l")
    nrm = cpu.estimate_normals(td, 0.04, 20)
    r = cpu.registration_icp(sd, td, 0.05, mode="p2l", tgt_normals=nrm, max_iteration=8)
    T2, f2, r2, it2, _ = twin.registration_icp(sd, td, 0.05, mode="p2l", tgt_normals=nrm, max_iteration=8)
    np.testing.assert_allclose(r.transformation, T2, atol=1e-10)
    assert r.iterations == it2


# ---------------------------------------------------------------- A.8 filters
def test_outlier_filters_match_twin():
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.normal(scale=0.05, size=(3000, 3)), rng.uniform(-1, 1, size=(60, 3))])
    np.testing.assert_array_equal(cpu.statistical_outlier_mask(pts, 45, 3.5), twin.statistical_outlier_mask(pts, 45, 3.5))
    np.testing.assert_array_equal(cpu.statistical_outlier_mask(pts, 80, 2.0), twin.statistical_outlier_mask(pts, 80, 2.0))
    np.testing.assert_array_equal(cpu.radius_outlier_mask(pts, 8, 0.023), twin.radius_outlier_mask(pts, 8, 0.023))


def test_transform_divides_by_w():
    T = np.eye(4)
    T[3, 3] = 2.0
    out = cpu.transform(np.array([[2.0, 4.0, 6.0]]), T)
    np.testing.assert_array_equal(out, [[1.0, 2.0, 3.0]])


def test_resident_cpu_map_is_the_twin_of_the_resident_sequence():
    """oracle.cpp ResidentMap (the like-for-like CPU baseline: voxel hash + stencil search, nothing rebuilt per
    scan) must give exactly what pipeline.ResidentSequence gives by re-running registration_icp (KD-tree) on
    the exported centroids: same iterations, bit-identical poses, fitness, rmse and map, for a map voxel equal
    to the gate (27-cell stencil) and for the reference default 2 cm / 5 cm (7^3 stencil)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import lidar_slam_b200  # noqa: F401  (registers the package alias)
    from lidar_slam_b200 import synth
    from oracle import pipeline

    scene = synth.make_corridor_scene(3, length=14.0)
    poses = synth.corridor_trajectory(scene, 4, seed=3, step=0.02)
    scans = [synth.render_scan(scene, p, seed=300 + k).numpy()[::4].copy() for k, p in enumerate(poses)]
    origin = np.array([-20.0, -20.0, -20.0])
    for map_voxel in (0.05, 0.02):
        a = pipeline.ResidentSequence(map_voxel, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30)
        b = cpu.ResidentMapCPU(map_voxel, origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30)
        assert a.process(scans[0]) is None and b.process(scans[0]) is None  # the first scan is the map
        for s in scans[1:]:
            ra, rb = a.process(s), b.process(s)
            assert ra.iterations == rb.iterations > 0
            assert ra.fitness == rb.fitness and ra.inlier_rmse == rb.inlier_rmse
            np.testing.assert_array_equal(ra.transformation, rb.transformation)
        ca, ka = a.map_points()
        cb, kb, nb = b.map_points()
        np.testing.assert_array_equal(ka, kb)
        np.testing.assert_array_equal(ca, cb)
        assert nb.sum() == sum(len(s) for s in scans)
        b.close()
