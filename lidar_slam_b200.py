"""Import shim: the package directory is named ``3d-lidar-slam_b200`` (not a valid
Python identifier), so ``import lidar_slam_b200`` loads it through importlib and
aliases it under this name."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("3d-lidar-slam_b200")
sys.modules[__name__] = _pkg
