"""Host-side mirror of the reference's processor plugin interface
(src/slam/scripts/point_cloud_processors/ + process_pc_manager.py) and the
B200 implementation that plugs into it."""
from .algorithm_enum import AlgorithmType  # noqa: F401
from .b200_icp_processor import B200ICPProcessor  # noqa: F401
from .data_transfer import DataTransfer  # noqa: F401
from .generic_point_cloud_processor import ProcessPointClouds  # noqa: F401
from .process_pc_manager import set_algorithm  # noqa: F401
