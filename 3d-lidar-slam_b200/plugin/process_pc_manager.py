"""Processor factory, mirroring ProcessPointCloudsHandler.set_algorithm
(src/slam/scripts/process_pc_manager.py:78-98) with the one added branch a maintainer
adds to select the B200 processor.  The reference's process loop
(process_pc_manager.py:35-57) is host control flow and stays as it is."""
from __future__ import annotations

from .algorithm_enum import AlgorithmType


def set_algorithm(algorithm: str, config_path: str, data_transfer, reset_event, **kwargs):
    try:
        algo = AlgorithmType[algorithm.upper()]
    except KeyError as exc:  # the reference lets the KeyError escape (process_pc_manager.py:80)
        raise ValueError(f"Invalid algorithm type: {algorithm}") from exc
    if algo == AlgorithmType.B200_ICP:
        from .b200_icp_processor import B200ICPProcessor

        return B200ICPProcessor(config_path=config_path, data_transfer=data_transfer, reset_event=reset_event,
                                **kwargs)
    if algo == AlgorithmType.ICP:
        raise ValueError("AlgorithmType.ICP is the reference's Open3D processor; it is not part of this package")
    if algo == AlgorithmType.DEEP_LEARNING:
        return None  # vacant slot in the reference too (process_pc_manager.py:91-92)
    raise ValueError(f"Unsupported algorithm: {algo}")
