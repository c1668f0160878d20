import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def b2():
    import lidar_slam_b200

    return lidar_slam_b200


@pytest.fixture(scope="session")
def engine(b2):
    eng = b2.Engine(0)
    yield eng
    eng.close()


_DRIVERS = ["group", "sweep"] + (["sweep4"] if os.environ.get("B2_EXPERIMENTS", "0") not in ("", "0") else [])


@pytest.fixture(params=_DRIVERS)
def driver(request, engine):
    """Run a test under both ICP-step kernels of the pair / slab grids (include/b2slam.h B2_ICP_DRIVER_*):
    the 4-lane-group kernel and the sweep kernel must give identical results.  The four-lanes-per-query sweep
    variant is an experiment that only exists in B2_EXPERIMENTS=1 builds."""
    if request.param == "sweep4" and b"+experiments" not in engine.lib.b2_version():
        pytest.skip("SWEEP4 driver: experiment, not in the default library (B2_EXPERIMENTS=1 builds it)")
    engine.set_icp_driver(request.param)
    yield request.param
    engine.set_icp_driver("auto")


@pytest.fixture(params=["persistent", "chained"])
def dense_driver(request, engine):
    """Run a test under both launch schemes of a registration against the brick-major resident map
    (include/b2slam.h B2_DENSE_DRIVER_*): one persistent launch, or one launch per ICP step.  Same kernel body,
    identical results."""
    import gc

    gc.collect()  # (the persistent launch is only taken while this engine is the only live handle on the device)
    engine.set_dense_driver(request.param)
    yield request.param
    engine.set_dense_driver("persistent")


@pytest.fixture(scope="session")
def pair0():
    """Config-1-shaped pair (seed 0): float32 source / target scans and the true relative pose."""
    from lidar_slam_b200 import synth

    src, tgt, T_true = synth.make_pair(0)
    return src.numpy(), tgt.numpy(), T_true


def rot_angle(Ra, Rb) -> float:
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def assert_transform_close(Ta, Tb, rad=1e-4, metres=1e-4):
    """north_star tolerance: 1e-4 rad, 1e-4 m."""
    Ta, Tb = np.asarray(Ta).reshape(4, 4), np.asarray(Tb).reshape(4, 4)
    ang = rot_angle(Ta[:3, :3], Tb[:3, :3])
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    assert ang < rad, f"rotation differs by {ang} rad"
    assert dt < metres, f"translation differs by {dt} m"
