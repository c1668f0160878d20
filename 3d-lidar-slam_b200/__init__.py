"""B200-native scan registration (ICP) + global-map fusion hot path of
MatthewKazan/3D-lidar-slam, behind the reference's ``ProcessPointClouds`` plugin
interface.

Layout
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/b2slam.h)
  _lib.py      ctypes binding of libb2slam.so; torch tensors are device buffers only
  plugin/      host-side mirror of the reference plugin interface + B200ICPProcessor
  synth.py     synthetic iPhone-LiDAR-shaped scans (test / bench inputs)

The directory name is not a Python identifier; ``import lidar_slam_b200`` (the shim
at the repository root) or ``importlib.import_module("3d-lidar-slam_b200")`` loads it.
"""
from . import _lib  # noqa: F401
from ._lib import B2Error, Engine, IcpParams, IcpResult, P2L, P2P, make_params, results_to_numpy  # noqa: F401

__all__ = ["Engine", "IcpParams", "IcpResult", "B2Error", "P2P", "P2L", "make_params", "results_to_numpy"]
