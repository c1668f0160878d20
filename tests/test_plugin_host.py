"""Host-side logic of the plugin mirror (no GPU): projection semantics, factory, sharding."""
import multiprocessing
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, "3d-lidar-slam_b200", "config", "lidar_config.yaml")


class _Logger:
    def info(self, *_): pass
    def debug(self, *_): pass


class _Proc:
    """Concrete stand-in so the abstract base's host logic can be exercised."""

    def __new__(cls, *a, **k):
        from lidar_slam_b200.plugin import ProcessPointClouds

        class P(ProcessPointClouds):
            def construct_global_map(self, points):
                self.seen = points

            def downsample_global_map(self):
                return self.global_map

        return P(*a, **k)


def test_load_config_rescales_intrinsics(b2):
    from lidar_slam_b200 import synth
    from lidar_slam_b200.plugin import DataTransfer

    p = _Proc(CONFIG, multiprocessing.Event(), DataTransfer(), _Logger())
    fx, fy, cx, cy = synth.intrinsics()
    assert (p.WIDTH, p.HEIGHT) == (256, 192)
    assert p.fx == pytest.approx(179.06465, abs=1e-4) and p.fx == fx and p.fy == fy
    assert p.cx == pytest.approx(128.74856, abs=1e-4) and p.cx == cx
    assert p.cy == pytest.approx(95.72725, abs=1e-4) and p.cy == cy
    assert p.point_clouds_in_map == 0 and p.global_map is None
    np.testing.assert_array_equal(p.previous_transformation[-1], np.eye(4))


def test_process_projects_and_counts(b2):
    from lidar_slam_b200.plugin import DataTransfer

    dt = DataTransfer()
    p = _Proc(CONFIG, multiprocessing.Event(), dt, _Logger())
    rec = np.array([[10.0, 20.0, 1.5], [128.0, 96.0, 2.0]], dtype=np.float32)
    dt.pixel_depth_map_queue.put(rec)
    p.process()
    assert p.point_clouds_in_map == 1
    assert p.seen.dtype == np.float32 and p.seen.shape == (2, 3)
    x = (10.0 - p.cx) * 1.5 / p.fx
    assert p.seen[0, 0] == pytest.approx(x, rel=1e-6) and p.seen[1, 2] == 2.0


def test_vectorised_projection_matches_the_loop(b2):
    """B200ICPProcessor.project_pixel_to_3d (vectorised) == the reference's per-point loop, in both
    NumPy scalar-promotion regimes."""
    from lidar_slam_b200 import synth
    from lidar_slam_b200.plugin.b200_icp_processor import B200ICPProcessor
    import torch

    rng = np.random.default_rng(0)
    rec = np.stack([rng.integers(0, 256, 500), rng.integers(0, 192, 500), rng.uniform(0.3, 5, 500)], 1).astype(np.float32)
    fx, fy, cx, cy = synth.intrinsics()
    loop = np.zeros((500, 3), dtype=np.float32)
    for i, (x, y, z) in enumerate(rec):  # scalars are np.float32: promotion follows the installed NumPy
        loop[i] = [(x - cx) * z / fx, (y - cy) * z / fy, z]
    numpy2 = int(np.__version__.split(".")[0]) >= 2
    fake = type("F", (), {"fx": fx, "fy": fy, "cx": cx, "cy": cy, "_numpy2": numpy2})()
    got = B200ICPProcessor.project_pixel_to_3d(fake, rec)
    np.testing.assert_array_equal(got, loop)
    ref = synth.project_records(torch.from_numpy(rec), mode="f32" if numpy2 else "f64").numpy()
    np.testing.assert_array_equal(got, ref)
    fake._numpy2 = not numpy2
    other = B200ICPProcessor.project_pixel_to_3d(fake, rec)
    np.testing.assert_allclose(other, loop, rtol=0, atol=2e-6)  # the two regimes differ by float32 rounding only


def test_structured_records_are_accepted(b2):
    from lidar_slam_b200.plugin.b200_icp_processor import _records_to_array

    st = np.zeros(3, dtype=[("x", np.float32), ("y", np.float32), ("z", np.float32)])
    st["x"], st["y"], st["z"] = [1, 2, 3], [4, 5, 6], [7, 8, 9]
    np.testing.assert_array_equal(_records_to_array(st), [[1, 4, 7], [2, 5, 8], [3, 6, 9]])


def test_factory_and_enum(b2):
    from lidar_slam_b200.plugin import AlgorithmType, DataTransfer, set_algorithm

    assert AlgorithmType["ICP"].value == "ICP" and AlgorithmType["B200_ICP"].value == "B200ICP"
    with pytest.raises(ValueError):
        set_algorithm("nope", CONFIG, DataTransfer(), multiprocessing.Event())
    assert set_algorithm("deep_learning", CONFIG, DataTransfer(), multiprocessing.Event()) is None


def test_processor_fails_loudly_without_gpu(b2):
    import torch
    from lidar_slam_b200.plugin import DataTransfer, set_algorithm

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        set_algorithm("b200_icp", CONFIG, DataTransfer(), multiprocessing.Event())


def test_shard_range_partitions_exactly(b2):
    from lidar_slam_b200.batch import shard_range

    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_synthetic_scan_shape(pair0):
    src, tgt, T = pair0
    assert src.shape == (49152, 3) and src.dtype == np.float32
    assert tgt.shape == (49152, 3)
    assert 0.3 <= src[:, 2].min() and src[:, 2].max() <= 5.0
    np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
    assert 0.03 <= np.linalg.norm(T[:3, 3]) <= 0.0501
