"""Algorithm selector, mirroring src/slam/scripts/algorithm_enum.py:3-5 with one
added member for the B200 processor (the reference looks members up by upper-cased
name: process_pc_manager.py:80)."""
from enum import Enum


class AlgorithmType(Enum):
    ICP = "ICP"
    DEEP_LEARNING = "DICP"
    B200_ICP = "B200ICP"
