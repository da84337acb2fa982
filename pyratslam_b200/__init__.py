"""pyratslam_b200 -- B200-native pose-cell network and view-template matcher (pyratslam's hot path).

Python host classes mirroring the reference's surface; every computation runs in hand-written
sm_100a CUDA kernels behind the C ABI of ``include/pyratslam_b200.h``.  No CPU fallback.
"""
from .experience_map import ExperienceMap  # noqa: F401
from .posecell_network import PoseCellEnsemble, PoseCellNetwork, PosecellNetwork  # noqa: F401
from .view_templates import ShardedViewTemplates, ViewTemplate, ViewTemplates  # noqa: F401

__version__ = "0.1.0"
